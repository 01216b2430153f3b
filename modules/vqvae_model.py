"""Drop-in for /root/reference/modules/vqvae_model.py."""
from mage_b200.model import VectorQuantizedVAE  # noqa: F401
