"""Counts the tensor-core / TMEM / TMA / mbarrier / cluster instructions in the SASS of the hot-path kernels
(profiles/r01f_sass_evidence.txt).  Usage: cd reconvat_b200/csrc && python ../../tools/sass_evidence.py"""
import subprocess, re, collections, sys
out = []
pats = [("rvb_stft_gemm.o", "fold2c_pair_kernelILi4"), ("rvb_stft_gemm.o", "fold2_pair_kernelE"), ("rvb_stft_gemm.o", "fold_pair_kernelINS_8MelTable"),
        ("rvb_frontend.o", "logmel_normalise_cluster_kernelILi512"), ("rvb_frontend.o", "fold_split_f16_kernelIsLb1ELb1")]
keep = re.compile(r"^(UTCHMMA|UTCQMMA|UTCBAR|UTMALDG|UBLKCP|LDTM|SYNCS|UTCATOMSWS|UCGABAR|ATOMG|RED|MUFU|STS|LDS|ATOMS|BAR|F2FP|I2F|STG|LDG|LDC|MEMBAR|ERRBAR|CCTL)")
for obj, pat in pats:
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, cnt, total = None, collections.Counter(), 0
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1) if pat in m.group(1) else None
            if cur: name = cur
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                total += 1
                op = m.group(1)
                if keep.match(op):
                    op = re.sub(r"\.(E|STRONG|GPU|FTZ|RN|64|128|U8|U16|S16|F32|CONSTANT|SYS)\b", "", op)
                    cnt[op] += 1
    out.append("## %s  (%d SASS instructions)" % (name, total))
    for op, c in sorted(cnt.items(), key=lambda kv: (-kv[1], kv[0])):
        out.append("  %5d  %s" % (c, op))
    out.append("")
print("\n".join(out))
