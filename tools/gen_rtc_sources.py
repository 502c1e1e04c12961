#!/usr/bin/env python3
"""gen_rtc_sources.py <out> <files...>: the headers of the run-time compiled Sobol kernel as ONE C++ raw string literal
(tiktak_b200/csrc/rtc.cu includes the result).  Quoted #include lines are dropped: the files are pasted in dependency order."""
import sys
out, files = sys.argv[1], sys.argv[2:]
parts = []
for f in files:
    t = open(f).read()
    assert ')TKRTC"' not in t
    t = "\n".join(l for l in t.split("\n") if not l.strip().startswith('#include "'))
    parts.append("/* ---- %s ---- */\n%s\n" % (f.split("/")[-1], t))
src = "\n".join(parts)
chunks = [src[i:i + 12000] for i in range(0, len(src), 12000)]  # (some compilers cap the length of one literal)
open(out, "w").write("\n".join('R"TKRTC(' + c + ')TKRTC"' for c in chunks) + "\n")
