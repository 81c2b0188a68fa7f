"""TEST INFRASTRUCTURE ONLY -- empty stand-in (import only)."""
