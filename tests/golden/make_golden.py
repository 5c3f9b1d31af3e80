#!/usr/bin/env python
"""Regenerate tests/golden/ from the REFERENCE itself (run in the build container only).

Every fixture is a capture of real mm_chain_dp calls made by the reference CLI built in place from
/root/reference (oracle/_ref/minimap2-sw, software chaining, `make -C oracle ref`), recorded by
oracle/dump_shim.c: inputs a[], the reference's f/p/v after the DP fill, and its final u[]/b[].
PAF md5s of the same runs go to paf_md5.json.  Inputs are defined in cases.py.
Usage:  python tests/golden/make_golden.py
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

CLI = os.path.join(cases.ROOT, "oracle", "_ref", "minimap2-sw")


def run(name, args, md5s):
    with tempfile.TemporaryDirectory() as td:
        dump = os.path.join(td, "d.bin")
        env = dict(os.environ, MM2_DUMP=dump)
        paf = subprocess.run([CLI, "-t", "1"] + args, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        md5s[name] = dict(args=[os.path.basename(a) if os.path.isabs(a) else a for a in args],
                          md5=hashlib.md5(paf).hexdigest(), lines=paf.count(b"\n"))
        with open(dump, "rb") as fi, gzip.GzipFile(os.path.join(HERE, name + ".dump.gz"), "wb", 9, mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
    print(name, md5s[name]["md5"], md5s[name]["lines"], os.path.getsize(os.path.join(HERE, name + ".dump.gz")))


def main():
    md5s = {}
    td = tempfile.mkdtemp()
    for name, args in cases.build_cases(td):
        run(name, args, md5s)
    shutil.rmtree(td)
    with open(os.path.join(HERE, "paf_md5.json"), "w") as fh:
        json.dump(md5s, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
