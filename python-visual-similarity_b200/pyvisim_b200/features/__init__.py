from ._features import SIFT, RootSIFT, Lambda, Descriptors, DeepConvFeature, FeatureExtractorBase

__all__ = ["SIFT", "RootSIFT", "Lambda", "Descriptors", "DeepConvFeature", "FeatureExtractorBase"]
