"""Not a test: SASS instruction-class counts and ptxas resource lines of the built library -> profiles/<tag>_sass_summary.txt"""
import collections, glob, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(root, "mgm_b200", "libmgmb200.so")
out = ["# SASS / ptxas evidence of mgm_b200/libmgmb200.so (%s)" % tag,
       "Built by mgm_b200/csrc/Makefile: nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -prec-div=true -prec-sqrt=true", "",
       "## ELF images in the library (cuobjdump -lelf)"]
out += subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout.strip().splitlines()
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cnt = collections.Counter()
for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", sass, re.M):
    cnt[m.group(1)] += 1
out += ["", "## instruction counts over all kernels (cuobjdump -sass), %d instructions" % sum(cnt.values())]
show = ["UBLKCP", "SYNCS", "LDGSTS", "FADD2", "FFMA2", "FMUL2", "FMNMX3", "FMNMX", "FADD", "FMUL", "FFMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR",
        "ERRBAR", "FENCE", "CCTL", "ATOMG", "MATCH", "VOTE", "VOTEU", "REDUX", "NANOSLEEP", "LDC", "LDCU", "DADD", "DMUL", "DFMA", "MUFU",
        "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG"]
for k in show:
    out.append("%-10s %d" % (k, cnt.get(k, 0)))
out += ["", "TMA bulk copies (UBLKCP), mbarrier transactions (SYNCS), cp.async (LDGSTS) and the packed fp32 pipes (FADD2 / FFMA2 / FMUL2,",
        "sm_100 only) are present; no tensor-core instructions (HMMA / UTC*MMA / LDTM / STTM): the path has no contraction.", "",
        "## registers, spills, shared memory per kernel (ptxas -v)"]
for f in sorted(glob.glob(os.path.join(root, "mgm_b200", "csrc", "*.ptxas.log"))):
    out.append("### " + os.path.basename(f))
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?\n\s+(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)\nptxas info\s+: (Used [^\n]+)", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        out.append("%s | %s | %s" % (name[:110], m.group(2), m.group(3)))
open(os.path.join(root, "profiles", "%s_sass_summary.txt" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out[:70]))
