"""Stages the reference's bundled INPUT DATA (no source code) for the GPU box, where /root/reference does not exist.

  * test_data/refs/*.fasta (the four phage assemblies of test.sh:3,11; 128 KB) -> tests/golden/phage/*.fasta.gz, committed;
  * refs/*.fasta (16 bacterial assemblies, 45 MB) -> oracle/_ref/refs/*.fasta.gz: git-ignored like every other file
    under oracle/_ref/, but not gpurun-ignored, so it travels with the snapshot.  __graft_entry__.build() calls this
    whenever the reference checkout is present.

Byte-identical copies (gzip, mtime 0 so the files are reproducible)."""
import glob
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFERENCE = "/root/reference"


def _stage(src, dst):
    data = open(src, "rb").read()
    if os.path.exists(dst):
        try:
            if gzip.open(dst, "rb").read() == data:
                return False
        except OSError:
            pass
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "wb") as f, gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=6) as g:
        g.write(data)
    return True


def stage(phage=True, genomes=True):
    if not os.path.isdir(REFERENCE):
        return 0
    n = 0
    if phage:
        for p in sorted(glob.glob(os.path.join(REFERENCE, "test_data/refs/*.fasta"))):
            n += _stage(p, os.path.join(ROOT, "tests/golden/phage", os.path.basename(p) + ".gz"))
    if genomes:
        for p in sorted(glob.glob(os.path.join(REFERENCE, "refs/*.fasta"))):
            n += _stage(p, os.path.join(ROOT, "oracle/_ref/refs", os.path.basename(p) + ".gz"))
    return n


if __name__ == "__main__":
    print(f"staged {stage()} file(s)", file=sys.stderr)
