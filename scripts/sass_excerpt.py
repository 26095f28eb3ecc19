"""Per-kernel SASS evidence of the in-tree libmpo_b200.so (no GPU needed):
   python scripts/sass_excerpt.py > profiles/<name>.txt
Counts, per kernel, the mnemonics that prove the hardware path (see the header it prints)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "multimodal-path-omic_b200", "libmpo_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "MUFU", "RED", "UCGABAR_ARV"]
print("# SASS evidence of the shipped multimodal-path-omic_b200/libmpo_b200.so (cuobjdump -sass, sm_100a); made by scripts/sass_excerpt.py.")
print("# Per kernel: instruction count and the count of the mnemonics that prove the hardware path:")
print("#   UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = TMA tensor load / store,")
print("#   UBLKCP = cp.async.bulk (1-D bulk copy), SYNCS = mbarrier, UTCBAR = tcgen05.commit, FFMA2 = packed fp32 FMA (tail).")
print()
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for i, blk in enumerate(blocks):
    body = blk.split("\n", 1)[1] if "\n" in blk else ""
    ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
    cnt = collections.Counter()
    for o in ops:
        base = o.split(".")[0]
        for w in WANT:
            if base == w:
                cnt[w] += 1
    name = names[i] if i < len(names) else blk.split("\n")[0]
    # drop the argument list: the last top-level parenthesised group of the demangled name
    depth, cut = 0, len(name)
    for j in range(len(name) - 1, -1, -1):
        if name[j] == ")":
            depth += 1
        elif name[j] == "(":
            depth -= 1
            if depth == 0:
                cut = j
                break
    name = name[:cut].replace("(bool)1", "true").replace("(bool)0", "false")
    tags = "  ".join("%s=%d" % (w, cnt[w]) for w in WANT if cnt[w])
    print("%-74s %5d instr  %s" % (name[:74], len(ops), tags))
