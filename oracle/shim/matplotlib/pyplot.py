"""TEST INFRASTRUCTURE ONLY -- empty stand-in for matplotlib.pyplot (import only)."""
