"""Opcode mix of every kernel in libb200moby.so (cuobjdump -sass), for profiles/: which pipes the code uses (DFMA / DADD /
DMUL FP64, LDS / STS shared memory, LDL / STL local memory, CREDUX warp reductions, MUFU.RCP64H division seeds, SHFL) and that
no tensor-core or TMA instruction is in it (the contractions on this path are 6-wide)."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "moby_b200/libb200moby.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, mix = None, collections.OrderedDict()
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        mix[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", l)
    if m and kern:
        op = m.group(1)
        if op == "MUFU":
            op += m.group(2)
        mix[kern][op] += 1
want = ("DFMA", "DADD", "DMUL", "DSETP", "LDS", "STS", "LDL", "STL", "LDG", "STG", "ATOMG", "CREDUX", "SHFL", "MUFU.RCP64H", "BRX", "BAR", "WARPSYNC", "HMMA", "DMMA", "UTCMMA", "UTMALDG", "UBLKCP")
print(f"{'kernel':58s} {'instr':>8s} " + " ".join(f"{w:>7s}" for w in want))
for k, c in mix.items():
    print(f"{k[:58]:58s} {sum(c.values()):8d} " + " ".join(f"{c.get(w, 0):7d}" for w in want))
