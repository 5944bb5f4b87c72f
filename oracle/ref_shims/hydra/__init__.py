"""TEST INFRASTRUCTURE ONLY -- the one hydra entry point the reference's model file uses (hydra is not in this image)."""
