"""Read an ncu report (source page, SASS) and split the kernel into regions of equal execution count:
for each region the share of the warp-state samples, of the executed instructions, and the leading stall
reasons.  Usage: python tools/ncu_regions.py gpurun_out/prof_dtmf.ncu-rep > profiles/rNN_ncu_dtmf_regions.txt"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    print(rows[0][0], rows[0][1] if len(rows[0]) > 1 else "")
    hdr = rows[1]
    data = rows[2:]
    isrc = hdr.index("Source")
    isamp = hdr.index("# Samples")
    iex = hdr.index("Instructions Executed")
    st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data)
    totex = sum(int(r[iex]) for r in data)
    print("instructions in kernel %d, samples %d, warp-instructions executed %d" % (len(data), tot, totex))
    regions = []
    cur = None
    for k, r in enumerate(data):
        ex = int(r[iex])
        if cur is None or abs(ex - cur[0]) > 0.02 * max(ex, cur[0], 1):
            cur = [ex, k, k, 0, 0]
            regions.append(cur)
        cur[2] = k
        cur[3] += int(r[isamp])
        cur[4] += ex
    print("%-13s %5s %11s %8s %8s  %-34s %s" % ("sass rows", "n", "exec/instr", "samples", "instr", "first instruction", "stalls"))
    for ex, a, b, s, e in regions:
        if s < tot * 0.004 and e < totex * 0.004:
            continue
        acc = {}
        for r in data[a:b + 1]:
            for i in st:
                v = int(r[i]) if r[i] not in ("", "-") else 0
                acc[hdr[i][6:]] = acc.get(hdr[i][6:], 0) + v
        t = max(sum(acc.values()), 1)
        top = ", ".join("%s %.0f%%" % (k, 100.0 * v / t) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:4])
        print("%5d-%-7d %5d %11d %7.1f%% %7.1f%%  %-34s %s" % (a, b, b - a + 1, ex, 100.0 * s / tot, 100.0 * e / totex,
                                                             data[a][isrc].strip()[:34], top))


if __name__ == "__main__":
    main()
