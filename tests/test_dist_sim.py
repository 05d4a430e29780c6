"""Config 5 (2-D block-cyclic distributed Cholesky / inverse / gradient): the multi-rank
driver with every rank inside one process (collectives = copies), on the SIMT emulator.
The identical driver, tables and kernels run over NCCL on real GPUs (tools/dist_check.py)."""
import pytest

from oracle import causalstump_oracle as o
from tests import _cases as cs


@pytest.mark.parametrize("kind,n,p,T,P,Q", [
    ("GP", 150, 3, 64, 2, 1), ("GP", 200, 4, 64, 2, 2), ("TP", 260, 3, 64, 2, 4), ("GP", 300, 2, 128, 2, 2),
    ("GP", 330, 2, 64, 1, 2), ("GP", 70, 1, 64, 2, 2),
])
def test_simulated_grid_matches_oracle(emulib, kind, n, p, T, P, Q):
    Pr = cs.make_problem(kind, n, p, weights=True)
    g, st, mu = o.eval_lml_grad(kind, Pr["y"], Pr["X"], Pr["z"], Pr["w"], Pr["par"])
    g2, st2, mu2 = emulib.dist_sim_eval(cs.kcode(kind), Pr["y"], Pr["X"], Pr["z"], Pr["w"], o.pack_params(Pr["par"]), P, Q, T)
    assert cs.rel(g2, o.pack_params(g), 1e-9) < cs.TOL_GRAD
    assert abs(st2[1] - st[1]) < cs.TOL_LML * abs(st[1]) and abs(st2[0] - st[0]) < 1e-8 * abs(st[0])
    assert abs(mu2 - mu) < 1e-9 * abs(mu)


def test_block_cyclic_ownership_is_a_partition():
    # every tile of the lower triangle is owned by exactly one rank of the grid
    for N, P, Q in ((7, 2, 1), (9, 2, 2), (16, 2, 4)):
        seen = {}
        for pr in range(P):
            for pc in range(Q):
                for J in range(N):
                    for I in range(J, N):
                        if I % P == pr and J % Q == pc:
                            assert (I, J) not in seen
                            seen[(I, J)] = (pr, pc)
        assert len(seen) == N * (N + 1) // 2
