from .pgtformer_arch import PGTFormer, TDCRQVAE3  # noqa: F401
