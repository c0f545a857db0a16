"""Scratch diagnostics for the tcgen05 kernel (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_tc_gpu import run_tc, merged, oracle_neg, data
from moma_b200 import _lib

lib = _lib.load()
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(128, 64, 1024), (128, 64, 128), (128, 64, 256)]
for (B, D, K) in shapes:
    for splits in (1,):
        q, queue = data(B, D, K, B + D + K)
        stats, Op, dbg, qb, kb = run_tc(lib, q, queue, 0.15, n_splits=splits, want_dbg=True)
        lse, Onorm, mx = merged(stats, Op)
        lse_o, O_o, mx_o = oracle_neg(qb, kb, 0.15)
        S = qb.float().cpu().numpy().astype(np.float64) @ kb.float().cpu().numpy().astype(np.float64).T
        BN = dbg.shape[1]
        print(f"== B{B} D{D} K{K} splits{splits}: S err {np.abs(dbg.cpu().numpy() - S[:, :BN]).max():.2e} "
              f"lse err {np.abs(lse - lse_o).max():.2e} mx err {np.abs(mx - mx_o).max():.2e}")
        err = np.linalg.norm(Onorm - O_o, axis=1) / np.linalg.norm(O_o, axis=1)
        print("   row err by 32-row warp:", [f"{err[i:i+32].max():.1e}" for i in range(0, B, 32)])
        cerr = np.linalg.norm(Onorm - O_o, axis=0) / np.linalg.norm(O_o, axis=0)
        print("   col err by 16-col group:", [f"{cerr[i:i+16].max():.1e}" for i in range(0, D, 16)])
        bad = np.argsort(-err)[:4]
        print("   worst rows", bad, err[bad])
        info = run_tc.info.cpu().numpy()
        for r in bad:
            print("     row", r, "rescale-calls", info[r, 0], "need-count", info[r, 1], "last f", info[r, 2], "last tile", info[r, 3])
        w = bad[0] // 32
        print("   warp", w, "per-lane need-count:", info[w*32:(w+1)*32, 1].astype(int).tolist())
        print("   warp", w, "per-lane err:", [f"{e:.2f}" for e in err[w*32:(w+1)*32]])

print("---- bounded scores (no rescale possible)")
for (B, D, K) in [(128, 64, 1024), (128, 64, 384)]:
    q, queue = data(B, D, K, B + D + K, qscale=0.05)
    stats, Op, dbg, qb, kb = run_tc(lib, q, queue, 0.15, n_splits=1, want_dbg=True)
    lse, Onorm, mx = merged(stats, Op); lse_o, O_o, mx_o = oracle_neg(qb, kb, 0.15)
    err = np.linalg.norm(Onorm - O_o, axis=1) / np.linalg.norm(O_o, axis=1)
    print(B, D, K, "row err by warp:", [f"{err[i:i+32].max():.1e}" for i in range(0, B, 32)])
print("---- forced rescale, NQ=1 (B<=128)")
for (B, D, K) in [(128, 128, 1024), (128, 64, 1024), (128, 256, 1024), (64, 128, 1024)]:
    rng = np.random.default_rng(1)
    q = O_.normalize(rng.standard_normal((B, D))).astype(np.float32) if False else None
    from oracle import moma_oracle as OO
    q = OO.normalize(rng.standard_normal((B, D))).astype(np.float32)
    queue = OO.normalize(rng.standard_normal((K, D))).astype(np.float32)
    direction = q.mean(axis=0); direction /= np.linalg.norm(direction)
    BN = 64 if D == 256 else 128
    for t in range(K // BN):
        queue[t * BN + 5] = direction * (0.5 + 0.8 * t) + 0.01 * queue[t * BN + 5]
    stats, Op, dbg, qb, kb = run_tc(lib, q, queue, 0.07, n_splits=1, want_dbg=True)
    lse, Onorm, mx = merged(stats, Op); lse_o, O_o, mx_o = oracle_neg(qb, kb, 0.07)
    err = np.linalg.norm(Onorm - O_o, axis=1) / np.linalg.norm(O_o, axis=1)
    print(B, D, K, "lse err", np.abs(lse - lse_o).max(), "row err by warp:", [f"{err[i:i+32].max():.1e}" for i in range(0, B, 32)])
