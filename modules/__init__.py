"""Same import paths as the reference (`target: modules.mage_model.MAGE` in config/*.yaml)."""
