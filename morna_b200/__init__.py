"""morna_b200 -- B200-native implementation of morna's data-parallel hot path
(feature-hashed index build + exact angular kNN search).  CUDA only."""
__all__ = ["MornaIndex", "MornaSearch", "go_index"]


def __getattr__(name):
    if name in ("MornaIndex", "go_index"):
        from . import index
        return getattr(index, name)
    if name == "MornaSearch":
        from .search import MornaSearch
        return MornaSearch
    raise AttributeError(name)
