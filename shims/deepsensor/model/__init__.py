from .convnp import ConvNP  # noqa: F401
