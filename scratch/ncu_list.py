import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    d = dict(zip(H, r))
    key = d["Kernel Name"][:60]
    a = agg.setdefault(key, collections.defaultdict(list))
    a[d["Metric Name"]].append(float(d["Metric Value"].replace(",", "")))
for k, a in agg.items():
    n = len(a["gpu__time_duration.sum"])
    t = sum(a["gpu__time_duration.sum"]) / n / 1e6
    rd = sum(a.get("dram__bytes_read.sum", [0])) / n / 1e9
    wr = sum(a.get("dram__bytes_write.sum", [0])) / n / 1e9
    print(f"{k:60s} n={n:3d} {t:9.3f} ms  rd {rd:7.3f} GB  wr {wr:7.3f} GB")
