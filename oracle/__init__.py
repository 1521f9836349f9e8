"""CPU oracle -- test infrastructure only (see oracle/mpc_oracle.h).  PARITY UNPINNED."""
