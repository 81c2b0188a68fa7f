"""TEST INFRASTRUCTURE ONLY -- empty stand-in for quimb.tensor (lib/tenn.py:23 imports it; only to_dense_quimb uses it)."""
