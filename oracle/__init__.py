"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see ttmpc_oracle.c header). Not imported by the product."""
