"""CPU oracle of the TSDF hot path — TEST INFRASTRUCTURE ONLY (see oracle/tsdf_oracle.c)."""
