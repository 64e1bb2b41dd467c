"""GRSS2013-shaped synthetic scene: 349x1905, 144-band uint16 HSI + 1 LiDAR, 15 classes."""
from hypelcnn_b200.loader.SyntheticDataLoader import SyntheticDataLoader


class SyntheticGRSS2013DataLoader(SyntheticDataLoader):
    pass
