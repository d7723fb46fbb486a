"""Synthetic Groth16 proving keys / assignments for the proof-assembly tests (shapes of
tachyon/zk/r1cs/groth16/proving_key.h and prove.h:54-61)."""
import numpy as np


def make_case(oracles, curve, n_full, n_pub, h_size, seed=0, h_query_size=None, blind=True):
    o1, o2 = oracles[curve], oracles[curve + "_g2"]
    n_wit = n_full - n_pub
    h_query_size = h_size if h_query_size is None else h_query_size
    pk = {
        "alpha_g1": o1.generate_points(seed + 1, 1)[0], "beta_g1": o1.generate_points(seed + 2, 1)[0],
        "delta_g1": o1.generate_points(seed + 3, 1)[0],
        "beta_g2": o2.generate_points(seed + 4, 1)[0], "delta_g2": o2.generate_points(seed + 5, 1)[0],
        "a_g1_query": o1.generate_points(seed + 6, n_full + 1),
        "b_g1_query": o1.generate_points(seed + 7, n_full + 1),
        "b_g2_query": o2.generate_points(seed + 8, n_full + 1),
        "h_g1_query": o1.generate_points(seed + 9, h_query_size),
        "l_g1_query": o1.generate_points(seed + 10, n_wit),
    }
    pk["b_g1_query"][3] = 0          # identity entries occur in real keys (unused wires)
    pk["b_g2_query"][3] = 0
    full = o1.generate_scalars(seed + 11, n_full, "witness")
    witness = full[n_pub:].copy()
    h = o1.generate_scalars(seed + 12, h_size, "uniform")
    rs = o1.generate_scalars(seed + 13, 2, "uniform")
    r, s = (rs[0], rs[1]) if blind else (np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64))
    return pk, r, s, h, witness, full
