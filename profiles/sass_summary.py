"""Static SASS instruction-class summary of every kernel in libfa_b200.so (cuobjdump -sass; sm_100a).
usage: python profiles/sass_summary.py [libfa_b200.so] > profiles/r2_sass_summary.txt
Static counts (instructions in the binary, not executed ones): they show WHICH hardware paths a kernel uses -- FP32 / FP64
pipes, the async-copy engine (UBLKCP = cp.async.bulk, SYNCS = mbarrier), warp collectives -- and that no tensor-core
(HMMA / UTCMMA) or library code is present; the executed mix is in the ncu captures."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "webspeechanalyzer_b200", "libfa_b200.so")
CLASSES = [
    ("fp32", r"^(FFMA|FMUL|FADD|FSEL|FSETP|FMNMX|FCHK|F2F|FRND)"),
    ("fp64", r"^(DFMA|DMUL|DADD|DSETP|F2F\.F64|I2F\.F64|F2I\.F64)"),
    ("mufu", r"^MUFU"),
    ("int", r"^(IMAD|IADD|IADD3|LOP3|SHF|LEA|ISETP|SEL|PRMT|POPC|FLO|BREV|IABS|IMNMX|VIADD|VIMNMX|I2F|F2I|I2I|PLOP3|P2R|R2P|MOV|CS2R|S2R|SGXT|BMSK|UIADD3|ULOP3|USHF|UMOV|UIMAD|ULEA|UISETP|USEL|UPRMT|UFLO|UPOPC|S2UR|R2UR|UP2UR|UR2UP|UPLOP3|UBMSK|UBREV|VOTEU|UF2FP|UI2F|UF2I|UVIADD|UVIMNMX|UIABS|USGXT|HFMA2|IDP|VABSDIFF)"),
    ("ld/st global", r"^(LDG|STG|LD\.|ST\.|LDC|LDCU|ULDC|CCTL|ATOMG|ATOM|RED|PREFETCH|LDL|STL|ERRBAR)"),
    ("ld/st shared", r"^(LDS|STS|ATOMS|LDSM|STSM)"),
    ("async copy (TMA / bulk / mbarrier)", r"^(UBLKCP|UTMALDG|UTMASTG|UBLKRED|UTMAREDG|UTMACMDFLUSH|SYNCS|LDGSTS|LDGDEPBAR|DEPBAR|UTMAPF|UCGABAR|ELECT)"),
    ("warp collectives", r"^(SHFL|VOTE|REDUX|MATCH|WARPSYNC|NANOSLEEP|CREDUX)"),
    ("barrier / fence", r"^(BAR|MEMBAR|FENCE|B2R|R2B)"),
    ("tensor core", r"^(HMMA|IMMA|DMMA|UTCMMA|UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTCCP|QGMMA|HGMMA)"),
    ("control", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET|BRX|JMP|NOP|YIELD|BREAK|BMOV|WARPSYNC|ACQBULK|KILL|BPT|RPCMOV|ENDCOLLECTIVE)"),
]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, counts, total = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern).split("(")[0]
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for name, pat in CLASSES:
            if re.match(pat, op):
                counts[kern][name] += 1
                break
        else:
            counts[kern]["other:" + op.split(".")[0]] += 1
print(f"# {os.path.relpath(lib, ROOT)}: static SASS instruction classes per kernel (sm_100a, cuobjdump -sass)")
any_tc = False
for k in sorted(counts, key=lambda k: -total[k]):
    c = counts[k]
    any_tc = any_tc or c["tensor core"] > 0
    main = ", ".join(f"{n} {c[n]}" for n, _ in CLASSES if c[n])
    other = ", ".join(f"{n[6:]} {v}" for n, v in c.items() if n.startswith("other:"))
    print(f"{k}\n    {total[k]} instructions: {main}" + (f"; other: {other}" if other else ""))
print(f"# tensor-core instructions anywhere: {'yes' if any_tc else 'none'} (no contraction on this path, by design); "
      f"kernels with the bulk-copy engine (UBLKCP/SYNCS): "
      + ", ".join(sorted(k for k in counts if counts[k]['async copy (TMA / bulk / mbarrier)'])))
