"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/ude_oracle.h)."""
