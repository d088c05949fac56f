"""ORACLE - test infrastructure only (see lf_oracle.py / nets.py headers). Never imported by the product path."""
