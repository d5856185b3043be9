from pytorch_lightning import LightningModule  # noqa: F401  (c_network.py:4)
