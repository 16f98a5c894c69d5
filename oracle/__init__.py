"""CPU oracle (test infrastructure). See lk_oracle.py. Not imported by the product package."""
