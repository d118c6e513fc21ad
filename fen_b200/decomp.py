"""Host-side slab decomposition logic (no GPU, no numpy on the hot path): FEN's own 2decomp layout with
``(prow, pcol) = (1, P)`` (src/grid.f90:125,168-173; test/small_test/fields/methods.f90:21).

These are the rules the CUDA library applies internally (context.cu / comm.cu / poisson.cu); they are
restated here so that launchers (bench.py, tests, a Fortran/MPI driver) can size and place per-rank data, and
so that the N > 1 host logic can be tested with ``gloo`` on machines without a GPU.
"""
from __future__ import annotations


def slab_bounds(nx: int, ny: int, nz: int, nranks: int, rank: int):
    """x-pencil bounds ``lo(3), hi(3)`` (1-based, inclusive) of `rank`: x and y whole, z in equal slabs."""
    if nranks < 1 or not 0 <= rank < nranks:
        raise ValueError("bad rank %d of %d" % (rank, nranks))
    if nz % nranks:
        raise ValueError("nz = %d is not divisible by the number of ranks %d" % (nz, nranks))
    nzl = nz // nranks
    return (1, 1, rank * nzl + 1), (nx, ny, (rank + 1) * nzl)


def zpencil_bounds(nx: int, ny: int, nz: int, nranks: int, rank: int):
    """Bounds of `rank` in the z-pencil layout of the Poisson solver's last-direction stage: z whole, y split
    (src/grid.f90:172-173 with prow = 1)."""
    if ny % nranks:
        raise ValueError("ny = %d is not divisible by the number of ranks %d" % (ny, nranks))
    nyl = ny // nranks
    return (1, rank * nyl + 1, 1), (nx, (rank + 1) * nyl, nz)


def z_neighbours(nranks: int, rank: int, periodic: bool):
    """(front, back) neighbour ranks of a slab, -1 at a non-periodic domain end (2decomp update_halo wraps
    around when the direction was declared periodic in decomp_2d_init, src/halo.f90:33)."""
    lo, hi = rank - 1, rank + 1
    if lo < 0:
        lo = nranks - 1 if periodic else -1
    if hi >= nranks:
        hi = 0 if periodic else -1
    return lo, hi


def transpose_blocks(ny: int, nz: int, nranks: int, rank: int):
    """Blocks `rank` sends in the y-slab -> z-pencil transpose (transpose_y_to_z, src/poisson.f90:982):
    list of ``(dest, (j_lo, j_hi), (k_lo, k_hi))``, 1-based inclusive global index ranges."""
    (_, _, k0), (_, _, k1) = slab_bounds(1, ny, nz, nranks, rank)
    out = []
    for dest in range(nranks):
        (_, j0, _), (_, j1, _) = zpencil_bounds(1, ny, nz, nranks, dest)
        out.append((dest, (j0, j1), (k0, k1)))
    return out


def alltoall_bytes_per_gpu(nx: int, ny: int, nz: int, nranks: int, x_periodic: bool = True) -> int:
    """Bytes one GPU sends over NVLink per transpose, the (P-1)/P off-rank share of its slab: the half spectrum
    (nx/2 + 1 complex per line) when x is periodic, the full width when x is a Neumann direction (the DCT variants
    npn / nnn carry real data in a full-width complex array, poisson.cu)."""
    width = nx // 2 + 1 if x_periodic else nx
    local = width * ny * (nz // nranks) * 16
    return local * (nranks - 1) // nranks


def scatter_schedule(nranks: int, rank: int, blk: int):
    """Order in which the row-copy transpose kernel (poisson.cu: k_a2a_scatter) visits (destination, index) pairs:
    consecutive blocks cycle over the destination ranks, starting one past the sender, so that at any moment every
    rank is storing to a different peer.  Returns the list of ``(dest, idx)`` in block order."""
    return [((rank + 1 + bx % nranks) % nranks, ((rank + 1 + bx % nranks) % nranks) * blk + bx // nranks)
            for bx in range(nranks * blk)]


def gather_handles(all_gather, blob: bytes, nranks: int) -> bytes:
    """Concatenation of every rank's exported handle in RANK ORDER, as fen_gpu_comm_connect expects.
    ``all_gather(obj) -> list`` is any collective (torch.distributed.all_gather_object, MPI, threads)."""
    parts = list(all_gather(blob))
    if len(parts) != nranks:
        raise ValueError("all_gather returned %d parts for %d ranks" % (len(parts), nranks))
    n = len(blob)
    for r, p in enumerate(parts):
        if len(p) != n:
            raise ValueError("rank %d exported %d bytes, expected %d" % (r, len(p), n))
    return b"".join(parts)


# ---- granule-blocked layouts of the slab transposes (fen_b200/csrc/slab_bulk.cuh) ---------------------------------
# A granule is 8 consecutive kx (128 bytes of complex fp64).  The arrays the transposes WRITE keep the transposed index
# next to the granule, so that what one rank receives from one tile is a single contiguous run.

GRANULE = 8


def zpencil_blocked_offset(g: int, k: int, jl: int, kxi: int, nz: int, nyl: int) -> int:
    """Element offset of (granule g, global plane k, local line jl, kx % 8) in the z-pencil array
    ``Cz[((g*nz + k)*nyl + jl)*8 + kxi]`` a rank holds after transpose_y_to_z."""
    return ((g * nz + k) * nyl + jl) * GRANULE + kxi


def yslab_blocked_offset(g: int, j: int, zl: int, kxi: int, ny: int, nzl: int) -> int:
    """Element offset of (granule g, global line j, local plane zl, kx % 8) in the y-slab array
    ``Cy[((g*ny + j)*nzl + zl)*8 + kxi]`` a rank holds after transpose_z_to_y."""
    return ((g * ny + j) * nzl + zl) * GRANULE + kxi


def forward_runs(ny: int, nz: int, nranks: int, rank: int, g: int, zl: int):
    """What the y-forward kernel ships for its tile (granule g, local plane zl) -- all ny line elements of 8 kx: one run
    per destination, in the order the kernel issues them (rank + 1 first, own rank last).  Returns a list of
    ``(dest, j_first, n_elements, dest_offset)``: the elements j_first .. j_first + nyl - 1 (x 8 kx) of the tile, which
    are contiguous in the tile (``tile[j*8 + kxi]``) and land contiguously at ``dest_offset`` of dest's Cz."""
    nyl, nzl = ny // nranks, nz // nranks
    out = []
    for q in range(1, nranks + 1):
        dest = (rank + q) % nranks
        out.append((dest, dest * nyl, nyl * GRANULE, zpencil_blocked_offset(g, rank * nzl + zl, 0, 0, nz, nyl)))
    return out


def backward_runs(ny: int, nz: int, nranks: int, rank: int, g: int, jl: int):
    """The same for the z-solve kernel's tile (granule g, local line jl) -- all nz plane elements of 8 kx:
    ``(dest, k_first, n_elements, dest_offset)`` into dest's Cy."""
    nyl, nzl = ny // nranks, nz // nranks
    out = []
    for q in range(1, nranks + 1):
        dest = (rank + q) % nranks
        out.append((dest, dest * nzl, nzl * GRANULE, yslab_blocked_offset(g, rank * nyl + jl, 0, 0, ny, nzl)))
    return out


def pieces(n: int, nq: int):
    """Index ranges [lo, hi) of the nq pieces a chunked transpose cuts n planes / granules into (poisson.cu:
    solve_blocked); empty pieces are skipped."""
    return [(n * q // nq, n * (q + 1) // nq) for q in range(nq) if n * (q + 1) // nq > n * q // nq]
