"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel share table.
usage: python tools/summarize_launches.py launches.csv "<command that was profiled>" > profiles/rNN_launches_summary.txt"""
import collections, csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows:
    if r is hdr or r[ki] == "Kernel Name":
        continue
    try:
        ms = float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name).replace("t2h::gemm::", "gemm::")
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print("# ncu launch list summary: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none")
print("#   " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print(f"# {sum(cnt.values())} launches, {total:.2f} ms total device time (cold-cache, serialised: compare SHARES)")
mine = sum(v for k, v in tot.items() if k.startswith(("gemm::", "t2h::")))
print(f"# hand-written kernels (t2h:: / gemm::): {100 * mine / total:.1f}% of the device time")
print("  share        ms     n  kernel")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{100 * v / total:6.2f}%  {v:8.3f} {cnt[k]:5d}  {k[:110]}")
