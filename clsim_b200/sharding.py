"""Multi-GPU sharding of the step->photon path: no collective, steps shard by bunch.

The reference runs one converter per device and hands each bunch to the least-loaded one
(private/clsim/I3CLSimModule.cxx:611-638; server: one ROUTER socket feeding every converter,
I3CLSimServer.cxx:126-135).  Here: one process (or converter object) per GPU, each with its own
slice of the safe-prime multiplier table so that RNG streams never repeat across devices, and a
host-side merge keyed by the bunch identifier.
"""
import numpy as np

# Rows of the multiplier table reserved per device: >= 2 streams (creation, propagation) per
# resident thread of the fast kernel (148 SMs x 1024 threads x 2 = 303 104 on B200).
RNG_ROWS_PER_DEVICE = 327680


def rng_row_offset(rank):
    return int(rank) * RNG_ROWS_PER_DEVICE


# Behind the rows of 8 propagation engines: the streams of the two neighbours on the device.  Per device
# 131 072 rows: the photon -> MCPE converter (148 x 256 = 37 888 streams) and the step generator
# (2 x 148 x 256 = 75 776 streams).  The whole table has about 5.9 million rows.
AUX_ROWS_BASE = 8 * RNG_ROWS_PER_DEVICE
AUX_ROWS_PER_DEVICE = 131072
TOTAL_ROWS_RESERVED = AUX_ROWS_BASE + 8 * AUX_ROWS_PER_DEVICE


def mcpe_row_offset(rank):
    return AUX_ROWS_BASE + int(rank) * AUX_ROWS_PER_DEVICE


def stepgen_row_offset(rank):
    return AUX_ROWS_BASE + int(rank) * AUX_ROWS_PER_DEVICE + 49152


def shard_bunches(num_bunches, rank, world):
    """Round-robin assignment of bunch indices to ranks."""
    return list(range(int(rank), int(num_bunches), int(world)))


def split_steps(steps, world, granularity=1):
    """Split one step series into `world` contiguous shards whose sizes are multiples of
    `granularity` (the converter's workgroup size); the remainder goes to the last shard and is
    padded by the caller (steps.pad_to_granularity)."""
    n = len(steps)
    per = (n // world // granularity) * granularity
    bounds = [i * per for i in range(world)] + [n]
    return [steps[bounds[i]:bounds[i + 1]] for i in range(world)]


def merge_results(results):
    """results: iterable of (identifier, photons) from any device in any order
    -> dict identifier -> concatenated photons (the caller's frame bookkeeping key,
    I3CLSimClientModule.cxx:359-439)."""
    out = {}
    for ident, photons in results:
        out.setdefault(int(ident), []).append(photons)
    return {k: (np.concatenate(v) if len(v) > 1 else v[0]) for k, v in out.items()}
