"""CPU oracle of the EfficientSAM3 hot path.  TEST INFRASTRUCTURE ONLY -- see oracle/README.md."""
