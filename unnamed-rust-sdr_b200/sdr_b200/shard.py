"""Partitioning of the hot path across the GPUs of one box.  No collective is needed
on the data path: FFT batches and channels are independent, and a contiguous sample range only needs
the K-1 input samples to its left (the tap-length halo), which the owner of a range reads itself.

    unit_range        independent units (FFT blocks, channels): contiguous, near-equal split
    sample_range      a sample stream cut on multiples of `align` (the decimation D, so that the
                      kept-output indexing (j+1)D-1 of adapters/mod.rs:30-37 is shard invariant),
                      plus the left halo each shard must read
"""


def unit_range(n_units, world, rank):
    lo = n_units * rank // world
    hi = n_units * (rank + 1) // world
    return lo, hi


def sample_range(n_samples, world, rank, halo=0, align=1):
    """returns (lo, hi, halo_lo): the shard owns inputs [lo, hi) and must also read [halo_lo, lo).
    All cuts are multiples of `align`; the last shard takes the remainder."""
    groups = n_samples // align
    glo, ghi = unit_range(groups, world, rank)
    lo, hi = glo * align, ghi * align
    if rank == world - 1:
        hi = n_samples
    halo_lo = max(0, lo - halo)
    return lo, hi, halo_lo


def output_range(n_samples, world, rank, align=1):
    """kept-output indices [olo, ohi) produced by shard `rank` for decimation `align`"""
    lo, hi, _ = sample_range(n_samples, world, rank, 0, align)
    return lo // align, (hi // align)
