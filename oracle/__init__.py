"""Test infrastructure: CPU oracle of the P1 assembly hot path.  See p1_oracle.py."""
