"""neko_b200: B200-native implementation of NEKO's Gato training hot path
(GatoPolicy.forward + backward + masked loss) behind the reference's own GatoPolicy / train.py API."""
__version__ = "0.1.0"
