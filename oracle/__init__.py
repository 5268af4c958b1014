"""CPU oracle for the retrieval-evaluation hot path - TEST INFRASTRUCTURE, never imported by the product.

See `oracle/cmh_oracle.py` for the parity status of each function."""
