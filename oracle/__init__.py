"""CPU oracles for the FIRE hot path.  TEST INFRASTRUCTURE ONLY: nothing under fire_b200 may import this."""
