"""trajopt_b200 — B200-native batched SQP trajectory optimizer (hot path of tesseract-robotics/trajopt)."""
