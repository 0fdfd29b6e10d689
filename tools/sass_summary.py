"""cuobjdump -sass of one kernel -> opcode histogram + excerpts around the instructions that prove the hardware path
(UTCHMMA / LDTM / STTM = tcgen05 + TMEM, UBLKCP / UTMALDG = bulk copies (TMA), HMMA, FFMA2, SYNCS = mbarrier).

  python tools/sass_summary.py easyhybrid.jl_b200/_build/eh_var_tc.o 'k_epoch.*EngTc.*Li1ELi1ELb1E' > profiles/r2_sass_k_epoch_tc.txt"""
import collections, re, subprocess, sys

obj, pat = sys.argv[1], re.compile(sys.argv[2])
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs, cur = [], None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = [m.group(1), []]
        funcs.append(cur)
    elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        cur[1].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
sel = [f for f in funcs if pat.search(f[0])]
if not sel:
    sys.exit("no kernel matches; candidates:\n" + "\n".join(f[0] for f in funcs))
name, ins = sel[nth]
demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
print(f"# {obj}\n# {demangled}\n# {len(ins)} SASS instructions")
ops = collections.Counter()
for l in ins:
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", l)
    if m:
        op = m.group(2)
        base = op.split(".")[0]
        ops[base if base not in ("LDS", "STS", "LDL", "STL", "HMMA", "UTCHMMA", "LDTM", "STTM", "UBLKCP", "SYNCS", "BAR") else op] += 1
print("\n## opcode histogram (static)")
for op, c in ops.most_common(60):
    print(f"{c:6d}  {op}")
marks = r"UTCHMMA|LDTM|STTM|UTCBAR|UBLKCP|UTMALDG|UTCATOMSWS|HMMA|SYNCS|FFMA2"
print("\n## first occurrences of the marker instructions (with one line of context)")
seen = collections.Counter()
for i, l in enumerate(ins):
    m = re.search(marks, l)
    if m and seen[m.group(0)] < 3:
        seen[m.group(0)] += 1
        for j in range(max(0, i - 1), min(len(ins), i + 2)):
            print(ins[j][:150])
        print("        ...")
