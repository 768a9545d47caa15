"""Sharding a batch of independent streams across the GPUs of one box (one process per GPU).

The QOA path has no cross-stream dependency (QoaEncoder / QoaDecoder hold only their own state, src/lib.rs:59-68,
:85-91), so multi-GPU is pure partitioning: every rank takes a contiguous range of streams, balanced by work, and no
data-path collective exists.  The only communication is the bench's barrier and the max / sum of timings and counts.
"""
import os


def partition_contiguous(costs, world_size):
    """Splits range(len(costs)) into `world_size` contiguous ranges whose summed costs are as even as a single
    left-to-right sweep can make them.  Returns [(begin, end), ...] (end exclusive); ranges may be empty.
    This is qoa_b200_partition of the C ABI: the split qoa_b200_mctx uses inside one process is the split the ranks
    of a torchrun job use (bench.py --scaling strong)."""
    from .engine import partition
    return partition([max(int(c), 0) for c in costs], world_size)


def decode_cost(qoa_len):
    """Decode work is proportional to the slice count, i.e. to the encoded length."""
    return max(int(qoa_len), 1)


def encode_cost(channels, samples):
    """Encode work is proportional to samples x channels (the 16-candidate search runs on every sample)."""
    return max(int(channels) * int(samples), 1)


def my_range(costs, rank=None, world_size=None):
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world_size = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else world_size
    return partition_contiguous(costs, world_size)[rank]
