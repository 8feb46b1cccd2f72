"""Device-side model step functions (fused pull/compute/push) for the ML apps."""
