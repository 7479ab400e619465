"""The cluster chain kernel (csrc/chain_tcgen05.cu) against the launch chain it replaces and against the oracle.

The chain kernel runs the same arithmetic in the same order (same tcgen05 K sequence, same row-wise code), so its
results must be BIT-IDENTICAL to the separate launches (NOVA_B200_CHAIN=0); parity with the CPU oracle is then
checked at the north-star tolerance (bf16: 2e-2 per step vs the fp32 oracle on the bf16-rounded weights).
"""

import os

import pytest
import torch

from gpu_util import cpu_sd, relmax
from oracle import loop as OL
from oracle import scheduler as OS

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2


def _head(depth, D, chain, T3=True):
    """A bf16 head whose library handle was created with NOVA_B200_CHAIN = chain (the flag is read at creation)."""
    import nova_pointcloud_b200 as nb

    old = os.environ.get("NOVA_B200_CHAIN")
    os.environ["NOVA_B200_CHAIN"] = "1" if chain else "0"
    os.environ["NOVA_B200_CHAIN_ROWS"] = "1000000"  # also above the row count where the library would switch back
    try:
        head = nb.synth.make_head(D, depth, dtype=torch.bfloat16, device="cuda", patch_size=1 if T3 else 2,
                                  image_dim=3 if T3 else 4)
        head.handle()
    finally:
        if old is None:
            del os.environ["NOVA_B200_CHAIN"]
        else:
            os.environ["NOVA_B200_CHAIN"] = old
    return head


def _inputs(B, N, D, n, seed=5, T=3):
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(B, N, T, generator=g)
    z = torch.randn(B, N, D, generator=g).bfloat16()
    order = torch.rand(B, N, generator=g).argsort(dim=1)
    ids = None if n is None else order[:, :n].unsqueeze(-1).contiguous()
    return noise, z, ids


@pytest.mark.parametrize("depth,D,B,N,n", [
    (6, 768, 32, 64, 1),      # 32 rows: the first sets of the cosine schedule
    (6, 768, 32, 64, 5),      # 160 rows: two row blocks, the second ragged
    (6, 768, 32, 64, 51),     # 1632 rows: the largest set of cfg2
    (6, 1024, 8, 80, 37),     # NOVA-0.6B width, 296 rows
    (6, 1536, 5, 40, 27),     # NOVA-1.4B width, 135 rows
    (3, 1280, 3, 50, 43),     # mlp_d3w1280, 129 rows (one row in the second block)
    (2, 1792, 2, 40, 35),     # widths beyond the registry's heads: 7 and 8 column groups per lane (the library
    (2, 2048, 3, 30, 23),     # accepts multiples of 256 up to 2048), 70 / 69 rows
    (2, 256, 3, 200, None),   # all tokens, 600 rows, smallest width
    (0, 512, 2, 9, 4),        # no blocks: embed -> final modulation -> head
])
def test_chain_kernel_is_bit_identical_to_the_launch_chain(depth, D, B, N, n):
    steps = 3
    ts, sig = OS.schedule(steps)
    noise, z, ids = _inputs(B, N, D, n)
    outs = []
    for chain in (False, True):
        head = _head(depth, D, chain)
        for _ in range(3):  # eager, graph capture, graph replay
            out = head.sample_tokens(noise.cuda(), z.cuda(), ts, sig, None if ids is None else ids.cuda())
        outs.append(out.cpu())
    assert torch.isfinite(outs[1]).all()
    assert torch.equal(outs[0], outs[1])


def test_chain_kernel_token_dim_16():
    """Registry-default token layout (patch 2 x 2 x 4 channels, T = 16) goes through the generic head/embed code."""
    steps = 2
    ts, sig = OS.schedule(steps)
    noise, z, ids = _inputs(4, 30, 768, 11, T=16)
    outs = []
    for chain in (False, True):
        head = _head(2, 768, chain, T3=False)
        outs.append(head.sample_tokens(noise.cuda(), z.cuda(), ts, sig, ids.cuda()).cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("D,B,N,n", [(768, 4, 96, 40), (1024, 3, 64, 50)])
def test_chain_kernel_per_step_matches_oracle(D, B, N, n):
    """Teacher-forced single Euler steps through the chain kernel vs the fp32 oracle on the bf16-rounded weights."""
    from oracle import head as OH

    head = _head(6, D, True)
    sd = cpu_sd(head)  # fp32 copies of the bf16-rounded weights
    noise, z, ids = _inputs(B, N, D, n, seed=11)
    traj = []
    OL.denoise(sd, z.float(), OH.unpatchify(noise, 1, 3, N, 1), num_steps=25, pred_ids=ids, trajectory=traj)
    ts, sig = OS.schedule(25)
    sel = ids.expand(-1, -1, 3)
    worst = 0.0
    for i in (0, 7, 24):
        x_t, _, x_next_ref = traj[i]
        out = head.sample_tokens(x_t.cuda(), z.cuda(), ts[i:i + 1], sig[i:i + 2], ids.cuda()).cpu()
        worst = max(worst, relmax(out.gather(1, sel), x_next_ref.gather(1, sel)))
    assert worst < BF16_TOL, worst
