"""Developer tool: per-kernel SASS opcode histogram of the built library (cuobjdump -sass) -> markdown table.
    python tools/sass_hist.py [lib] > profiles/rNN_sass_opcodes.md"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "pytorch-asr_b200/torch_asr/libctc_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur][m.group(2).split(".")[0]] += 1
keys = ["UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "LDGSTS", "ATOMS", "MUFU", "SHFL", "BAR", "LDS", "STS", "STG",
        "LDG", "FFMA", "FADD", "FMUL", "F2I", "I2F", "S2R", "BRA", "CCTL"]
print("| kernel | instr | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for f, c in funcs.items():
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    print("| `" + name[:90] + "` | " + str(sum(c.values())) + " | " + " | ".join(str(c.get(k, 0)) for k in keys) + " |")
