"""TEST INFRASTRUCTURE ONLY -- `from jax.config import config` (dQA_mps.py:27)."""


class _Config:
    def update(self, _key, _value):
        return None


config = _Config()
