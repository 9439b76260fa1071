"""TEST INFRASTRUCTURE ONLY: CPU restatements of the reference hot path (see hgru_oracle_np.py).
Never imported by the product package."""
