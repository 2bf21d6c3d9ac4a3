"""Fixture generator (run once in the build container, where /root/reference exists): the REAL ContentVec frame features the
reference ships as data -- logs/reference/ref_feats.npy (169 x 768) and logs/mute/extracted/mute.npy (149 x 768) -- copied
into tests/golden/contentvec_real_318.npz, so that parity tests can run on rows with the real distribution (a few dominant
dimensions, heavy tails) on a box where /root/reference does not exist.  Data only; no reference code is involved.

    python tests/golden/make_real_feats_fixture.py
"""
import hashlib
import os

import numpy as np

REF = "/root/reference"
SOURCES = ["logs/reference/ref_feats.npy", "logs/mute/extracted/mute.npy"]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "contentvec_real_318.npz")


def main():
    parts, meta = [], []
    for rel in SOURCES:
        a = np.load(os.path.join(REF, rel)).astype(np.float32)
        a = a.reshape(-1, a.shape[-1])
        assert a.shape[1] == 768 and np.isfinite(a).all()
        parts.append(a)
        meta.append(f"{rel}:{a.shape[0]}:{hashlib.sha256(a.tobytes()).hexdigest()[:16]}")
    rows = np.concatenate(parts, axis=0)
    np.savez_compressed(OUT, rows=rows, sources=np.array(meta))
    print(OUT, rows.shape, os.path.getsize(OUT), meta)


if __name__ == "__main__":
    main()
