#!/usr/bin/env python
"""Compare the device code (SASS, instruction text) of two commits' CUDA sources, kernel by kernel:
    python profiles/sass_diff.py <validated-commit> [<commit, default: working tree>]
Used to show that clean-ups after the last GPU call of a round did not change what the GPU runs."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin"]


def kernels(src_dir):
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for cu in sorted(f for f in os.listdir(src_dir) if f.endswith(".cu")):
            cubin = os.path.join(tmp, cu + ".cubin")
            subprocess.run(["nvcc"] + FLAGS + ["-o", cubin, cu], cwd=src_dir, check=True)
            sass = subprocess.run(["cuobjdump", "-sass", cubin], stdout=subprocess.PIPE, text=True, check=True).stdout
            cur = None
            for line in sass.splitlines():
                m = re.match(r"\s*Function : (\S+)", line)
                if m:
                    cur = out.setdefault(m.group(1), [])
                    continue
                m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
                if m and cur is not None:
                    cur.append(m.group(1))
    return out


def checkout(commit, tmp):
    path = os.path.join(tmp, commit)
    subprocess.run(["git", "worktree", "add", "-q", "--detach", path, commit], cwd=ROOT, check=True)
    return path


def main():
    a_commit = sys.argv[1]
    b_commit = sys.argv[2] if len(sys.argv) > 2 else None
    with tempfile.TemporaryDirectory() as tmp:
        trees = [checkout(a_commit, tmp)] + ([checkout(b_commit, tmp)] if b_commit else [])
        try:
            a = kernels(os.path.join(trees[0], "slr-sfs_b200", "csrc"))
            b = kernels(os.path.join(trees[1] if b_commit else ROOT, "slr-sfs_b200", "csrc"))
        finally:
            for t in trees:
                subprocess.run(["git", "worktree", "remove", "--force", t], cwd=ROOT)
    # template-parameter changes rename kernels: match leftovers by identical bodies
    same = [k for k in a if k in b and a[k] == b[k]]
    diff = [k for k in a if k in b and a[k] != b[k]]
    only_a = [k for k in a if k not in b]
    only_b = [k for k in b if k not in a]
    renamed = [(x, y) for x in only_a for y in only_b if a[x] == b[y]]
    print("identical: %d kernels" % len(same))
    for x, y in renamed:
        print("identical under a new name: %s -> %s" % (x, y))
    for k in diff:
        print("DIFFERENT:", k, len(a[k]), "->", len(b[k]), "instructions")
    for k in only_a:
        if k not in [x for x, _ in renamed]:
            print("removed:", k)
    for k in only_b:
        if k not in [y for _, y in renamed]:
            print("NEW:", k)
    return 1 if diff else 0


if __name__ == "__main__":
    sys.exit(main())
