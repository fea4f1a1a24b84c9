"""Prepare an A/B of two builds of the library for one short GPU call.

    python tools/make_ab.py <git-ref> [-DMACRO=VALUE ...]

builds vlapy_b200/lib/libvpfp_b200_base.so from the sources of <git-ref> (a worktree under /tmp; the working tree is
not touched) and the working tree into vlapy_b200/lib/libvpfp_b200.so -- or, when -D flags are given (a candidate
behind a macro), into vlapy_b200/lib/libvpfp_b200_cand.so, so that the product library never carries a candidate's
flags.  On the GPU box:

    python tools/time_libs.py vlapy_b200/lib/libvpfp_b200_base.so vlapy_b200/lib/libvpfp_b200[_cand].so   # e df/dv, one process
    VPFP_B200_LIB=$PWD/vlapy_b200/lib/libvpfp_b200_base.so python tools/time_ops.py 16384 16384 "<ops>"   # any operator
    python tools/time_ops.py 16384 16384 "<ops>"

(stage `ab:<ops>` of `tools/gpu_session.sh` runs exactly this; a call of this kind costs 30-80 s of GPU time.)  Also prints the
ptxas register / spill summary of the kernels whose numbers changed between the two builds -- the first thing to look
at before spending GPU time (a jump in spill bytes has so far always meant a slower kernel).  Delete the base library
before committing measurements: only libvpfp_b200.so is the product."""
import os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vlapy_b200 import _lib


def ptxas_summary(text):
    out, name = {}, None
    for line in text.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name and "spill" not in out.setdefault(name, {}):      # (later lines belong to device functions)
            out[name]["spill"] = (int(m.group(2)), int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m and name and "regs" not in out.setdefault(name, {}):
            out[name]["regs"] = int(m.group(1))
    return out


def build(src_root, so, defines):
    cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + _lib.NVCC_FLAGS + ["-Xptxas", "-v"] + defines + \
          ["-I", os.path.join(src_root, "include"), "-o", so, os.path.join(src_root, "vlapy_b200", "csrc", "vpfp_cuda.cu")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode:
        sys.exit(p.stderr[-3000:])
    return ptxas_summary(p.stderr)


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    ref, defines = sys.argv[1], sys.argv[2:]
    wt = "/tmp/vpfp_ab_worktree"
    subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", wt], capture_output=True)
    subprocess.check_call(["git", "-C", ROOT, "worktree", "add", "--detach", wt, ref], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    try:
        base = build(wt, os.path.join(ROOT, "vlapy_b200", "lib", "libvpfp_b200_base.so"), [])
    finally:
        subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", wt], capture_output=True)
    target = os.path.join(ROOT, "vlapy_b200", "lib", "libvpfp_b200_cand.so") if defines else _lib.SO
    new = build(ROOT, target, defines)
    for k in sorted(set(base) | set(new)):
        if base.get(k) != new.get(k):
            short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:100]
            print("%-100s  base %s  new %s" % (short, base.get(k), new.get(k)))
    print("base:", ref, "-> vlapy_b200/lib/libvpfp_b200_base.so ; new: working tree", " ".join(defines), "->",
          os.path.relpath(target, ROOT))


if __name__ == "__main__":
    main()
