"""CPU oracle for the rlrouting hot path — TEST INFRASTRUCTURE ONLY (see rlr_oracle.c)."""
