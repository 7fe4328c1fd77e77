"""CPU oracle: test infrastructure only (see assembly_oracle.py header)."""
