"""Part of the TensorFlow stand-in under oracle/tfshim (TEST INFRASTRUCTURE ONLY)."""
