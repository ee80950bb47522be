from .gato_policy import GatoPolicy  # noqa: F401
