"""Drop-in for /root/reference/modules/mage_model.py: the classes config/*.yaml address by dotted
path, implemented in mage_b200.model on the sm_100a kernels."""
from mage_b200.model import MAGE, FlatAxialDecoder, MAEncoder, TransformerTextEncoder  # noqa: F401
