"""simple-hrnet_b200: B200-native (sm_100a) engine for the simple-HRNet inference hot path.

Public surface (mirrors the reference, stefanopini/simple-HRNet):
    SimpleHRNet(c, nof_joints, checkpoint_path, ...).predict(image)      # SimpleHRNet.py:12,174
    B200Engine                                                            # the `self.model` seam (SimpleHRNet.py:143-147)
    HostPipeline                                                          # double-buffered host -> joints serving loop
    evaluation.get_max_preds / get_final_preds / flip_back / flip_average # misc/utils.py:19-29,125-182 on the GPU
The compute lives in libhrnet_b200.so (hand-written CUDA: tcgen05 implicit-GEMM convs fed by TMA
im2col, fused BN/ReLU/residual epilogues, exchange-unit fusion, argmax decode) behind the C ABI
declared in include/hrnet_b200.h.  There is no CPU or PyTorch fallback: without the library and a
B200 the engine raises.
"""
from ._lib import load_library, library_path, HrnetError  # noqa: F401
from .engine import B200Engine, HostPipeline, pack_state_dict, expected_state_dict_keys  # noqa: F401
from .api import SimpleHRNet  # noqa: F401
from .dist import shard_range, ShardedPredictor  # noqa: F401
from . import evaluation  # noqa: F401

__all__ = ["SimpleHRNet", "B200Engine", "HostPipeline", "pack_state_dict", "expected_state_dict_keys", "load_library",
           "library_path", "HrnetError", "shard_range", "ShardedPredictor", "evaluation"]
