"""pyvisim_b200 -- B200-native encode-and-compare path behind pyvisim's own API.

    from pyvisim_b200.encoders import VLADEncoder, FisherVectorEncoder, Pipeline, KMeansWeights, GMMWeights
    from pyvisim_b200.features import SIFT, RootSIFT, Lambda, Descriptors
    from pyvisim_b200.eval import retrieve_top_k_similar, top_k_map, top_k_accuracy

The compute path is ``lib/libpvs_b200.so`` (hand-written CUDA for sm_100a behind the C ABI
of ``include/pvs_b200.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
__all__ = ["encoders", "features", "eval", "retrieval"]
