"""Opcode histogram of the loops of a kernel's SASS (cuobjdump -sass -fun <name> file.o > k.sass):
python tools/sass_loops.py k.sass [min_len].  A loop = a backward branch; the body is the address
range it spans.  Used to count the instructions of the unrolled van-Herk block loop per k-mer."""
import collections, re, sys
ins = []
for ln in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 100
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_idx and i - addr_idx[tgt] >= minlen:
            loops.append((addr_idx[tgt], i))
print(f"{len(ins)} instructions, {len(loops)} loops of >= {minlen}")
for lo, hi in loops:
    ops = collections.Counter()
    for a, t in ins[lo:hi + 1]:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    alu = sum(ops[k] for k in ("LOP3", "SHF", "VIMNMX", "VIMNMX3", "PRMT", "ISETP", "IADD3", "VIADD", "SEL", "LEA", "IADD", "MOV", "PLOP3", "VIADDMNMX", "IABS", "POPC", "BREV", "FLO"))
    print(f"loop {ins[lo][0]:#06x}-{ins[hi][0]:#06x}: {hi - lo + 1} instr, ALU-pipe ~{alu}: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
