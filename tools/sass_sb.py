"""Decode scoreboard / stall control bits of a kernel's SASS (sm_100a).  usage: sass_sb.py OBJ KERNEL_SUBSTR [start end]"""
import re, subprocess, sys
obj, want = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
out, on = [], False
i = 0
while i < len(txt):
    l = txt[i]
    if "Function :" in l:
        on = want in l
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", l)
        if m and i + 1 < len(txt):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", txt[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                out.append((int(m.group(1), 16), re.sub(r"\s+", " ", m.group(2).strip()), (hi >> 41) & 0xf, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f))
                i += 2
                continue
    i += 1
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi_ = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
only = len(sys.argv) > 5
for a, ins, st, wb, rb, wait in out:
    if lo <= a < hi_ and (not only or wait or wb != 7 or "BRA" in ins or "BSSY" in ins or "BSYNC" in ins):
        print("%05x st%2d W%s R%s wait%s  %s" % (a, st, wb if wb != 7 else "-", rb if rb != 7 else "-", format(wait, "06b"), ins[:100]))
print(len(out), "instructions")
