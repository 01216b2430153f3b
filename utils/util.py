"""Drop-in for /root/reference/utils/util.py:45-63 (the config-addressed plugin mechanism)."""
from mage_b200.config import get_obj_from_str, instantiate_from_config  # noqa: F401
