"""TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/ref_ops.py). Never imported by ocn_b200."""
