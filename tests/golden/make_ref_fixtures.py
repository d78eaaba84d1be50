#!/usr/bin/env python
"""Golden outputs of THE REFERENCE'S OWN CODE for the edge-case fixture set.

Runs oracle/_ref/cfrk_ref_cpu (the reference's unmodified main.cu + fastaIO.h + kmer_main.cu +
kmer_kernel.cu compiled for the CPU by oracle/Makefile, SURVEY.md 8c "oracle #1") on every
(fixture, k, chunkSize) of tests/fixtures.py and records the sha256 of each output, plus the
output itself when it is small, in tests/golden/ref_shim/.  Run in the build container (needs
/root/reference); the manifest travels to the GPU box, the reference does not.

(The real reference binary on a GPU cannot serve here: its reader strcat()s into uninitialised
malloc memory (src/fastaIO.h:51-52) and on the B200 box every read picked up stale heap bytes,
see DESIGN.md "what the reference binary does on a B200".)
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures as fx  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "cfrk_ref_cpu")
OUT = os.path.join(HERE, "ref_shim")
KEEP_BYTES = 24 * 1024


def cases():
    for name, text, ks in fx.EDGE_SET:
        for k in ks:
            if name == "R_ragged" and k > 6:
                continue
            yield name, text, k, 8192
    for name, text, ks, chunks in fx.CHUNK_SET:
        for k in ks:
            for ch in chunks:
                yield name, text, k, ch
    yield "H_like_seq2", fx.fx_like_seq(710, 151), 5, 8192
    yield "H_like_seq2", fx.fx_like_seq(710, 151), 3, 100


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    tmp = tempfile.mkdtemp()
    for name, text, k, ch in cases():
        fa = os.path.join(tmp, name + ".fa")
        with open(fa, "w") as f:
            f.write(text)
        out = os.path.join(tmp, "out.cfrk")
        if os.path.exists(out):
            os.remove(out)
        r = subprocess.run([REF, fa, out, str(k), "12", str(ch)], capture_output=True, timeout=600)
        data = open(out, "rb").read() if os.path.exists(out) else b""
        key = f"{name}.k{k}.c{ch}"
        manifest[key] = {"fasta_sha256": hashlib.sha256(text.encode()).hexdigest(), "k": k, "chunk": ch,
                         "rc": r.returncode, "stdout_empty": r.stdout == b"",
                         "out_sha256": hashlib.sha256(data).hexdigest(), "out_bytes": len(data)}
        if len(data) <= KEEP_BYTES:
            with open(os.path.join(OUT, key + ".cfrk"), "wb") as f:
                f.write(data)
        print(key, r.returncode, len(data))
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
