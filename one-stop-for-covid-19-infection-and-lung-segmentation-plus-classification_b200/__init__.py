"""b200unet: B200-native U-Net train/infer engine behind the reference's Keras-shaped surface."""
__version__ = "0.1.0"
