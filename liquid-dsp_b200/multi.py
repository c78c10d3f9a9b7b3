"""Host-side sharding of the channelizer path across ranks (one process per GPU).

Two ways to shard:
  * independent streams -> `streams_of_rank`: contiguous blocks of streams per rank, no collective;
  * one long stream     -> `segment_of_rank` + `exchange_halo`: rank r owns a time segment whose first
    frame is even; the last `halo` input samples of rank r's segment are sent to rank r+1
    (point-to-point send/recv: NCCL over NVLink on GPUs, gloo in the CPU tests).

Nothing here computes samples; the tensors passed in are whatever backend the process group uses.
"""


def streams_of_rank(n_streams, rank, world):
    """contiguous block [lo, hi) of stream indices owned by `rank`"""
    base, rem = divmod(n_streams, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def segment_of_rank(n_frames_total, rank, world):
    """frames [lo, hi) of one long stream owned by `rank`; every boundary is an even frame so that each
    segment starts with the reference's flag = 0 (firpfbch2.c:57) and execute alignment holds"""
    per = -(-n_frames_total // world)           # ceil
    per += per & 1                              # even
    lo = min(rank * per, n_frames_total)
    hi = min(lo + per, n_frames_total)
    return lo, hi


def exchange_halo(dist, head, tail, rank, world):
    """send `tail` (last halo samples of this rank's segment) to rank+1, receive the left neighbour's
    tail into `head` (rank 0's head stays zero = the reference's reset state).  Returns the pending
    requests so the caller can run the interior frames before waiting."""
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, tail, rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, head, rank - 1))
    # one batched group (ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd): the send and the receive of a
    # middle rank progress together instead of being serialised as two independent operations
    return dist.batch_isend_irecv(ops) if ops else []
