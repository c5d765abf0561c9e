import os, sys, json
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from ntlink_b200 import Context, SeqBatch
import util
rng = np.random.default_rng(5)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
lens = []; left = n
while left > 0:
    L = min(left, int(rng.integers(2000, 60000))); lens.append(L); left -= L
offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
batch = SeqBatch(seq, offs, [f"s{i}" for i in range(len(lens))])
ctx = Context(0)
for S, k, w in [(128, 20, 10), (256, 20, 10), (256, 32, 100), (128, 15, 5)]:
    ctx.set_option("strip_len", S)
    ref = None
    bad = 0
    for it in range(12):
        sk = ctx.sketch(batch, k, w)
        cur = (sk.hash.copy(), sk.pos_strand.copy(), sk.seq_off.copy())
        if ref is None:
            ref = cur
            continue
        same = len(cur[0]) == len(ref[0]) and np.array_equal(cur[0], ref[0]) and np.array_equal(cur[1], ref[1]) and np.array_equal(cur[2], ref[2])
        if not same:
            bad += 1
            # locate
            dq = np.flatnonzero(np.diff(cur[2].astype(np.int64)) != np.diff(ref[2].astype(np.int64)))
            info = {"it": it, "len_cur": len(cur[0]), "len_ref": len(ref[0]), "seqs_with_diff_count": dq[:5].tolist()}
            if len(dq):
                q = int(dq[0]); a0, a1 = int(cur[2][q]), int(cur[2][q + 1]); b0, b1 = int(ref[2][q]), int(ref[2][q + 1])
                gp = (cur[1][a0:a1] & 0x7FFFFFFF).tolist(); rp = (ref[1][b0:b1] & 0x7FFFFFFF).tolist()
                info.update({"q": q, "seq_len": lens[q], "extra": sorted(set(gp) - set(rp))[:8], "missing": sorted(set(rp) - set(gp))[:8]})
            else:
                d = np.flatnonzero(cur[0] != ref[0])[:5].tolist() if len(cur[0]) == len(ref[0]) else []
                info["hash_diff_at"] = d
            print(json.dumps(info), flush=True)
    print(json.dumps({"S": S, "k": k, "w": w, "repeats": 11, "differing": bad}), flush=True)
ctx.close()
