"""gpu-wah_b200 -- B200-native WAH compress/decompress behind GPU-WAH's host entry points.

The directory name carries a hyphen (it is the project name), so it cannot be imported by
name; ``import gpu_wah_b200`` (the loader module at the repo root) registers this package
under that name.  Everything here is a thin host-side mirror of the C ABI declared in
``include/wah_b200.h``; the compute lives in ``lib/libwah_b200.so`` (hand-written sm_100a
CUDA).  There is no CPU fallback: if the library is missing, import raises.
"""
from .wah import (  # noqa: F401
    WAH_BLOCK1024,
    WAH_CANONICAL,
    WAH_STATUS_BAD_WORDS_MASK,
    WAH_STATUS_BATCH_LENGTH,
    WAH_STATUS_TIMEOUT,
    ShardRecord,
    WahError,
    Workspace,
    compress,
    compress_batch_device,
    compress_device,
    container_pack,
    container_unpack,
    decoded_words,
    decompress,
    decompress_batch_device,
    decompress_device,
    decoded_size_device,
    gen_clustered_device,
    gen_uniform_device,
    lib,
    lib_path,
    logical_device,
    popcount_device,
    WAH_OP_AND,
    WAH_OP_OR,
    WAH_OP_XOR,
    WAH_OP_ANDNOT,
    max_compressed_words,
    num_groups,
    shard_record_device,
    stitch_plan,
)
from . import mgpu  # noqa: F401

__all__ = [
    "WAH_BLOCK1024", "WAH_CANONICAL", "WAH_STATUS_BAD_WORDS_MASK", "WAH_STATUS_TIMEOUT", "WAH_STATUS_BATCH_LENGTH", "WahError", "Workspace", "compress", "decompress",
    "compress_device", "compress_batch_device", "decompress_device", "decompress_batch_device", "decoded_size_device",
    "num_groups", "max_compressed_words", "decoded_words", "gen_uniform_device",
    "gen_clustered_device", "shard_record_device", "stitch_plan", "container_pack", "container_unpack", "logical_device", "popcount_device",
    "WAH_OP_AND", "WAH_OP_OR", "WAH_OP_XOR", "WAH_OP_ANDNOT", "lib", "lib_path",
    "mgpu",
]
