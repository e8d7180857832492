"""CPU oracle of the CFP fusion path — test infrastructure only (see cfp_oracle.py)."""
