#!/usr/bin/env python3
"""Executed warp instructions and stall samples per SOURCE LINE of one kernel: joins the SASS rows of an ncu source page
(CSV made on the GPU box, gzip) with the line table of the same build (nvdisasm -g of the object's cubin), by instruction order.
Usage: tools/ncu_lines.py gpurun_out/x_full_source.csv.gz av1_base_b200/csrc/me_kernels.o hme_refine_kernel [source.cu]"""
import collections, csv, gzip, os, re, subprocess, sys, tempfile


def main():
    page, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
    lines, cur, on = [], None, False
    for l in dis:
        if l.startswith("//----") and ".text." in l:
            on = kern in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    ex, st = collections.Counter(), collections.Counter()
    i, take, done = 0, False, False
    for r in csv.reader(gzip.open(page, "rt")):
        if not r:
            continue
        if r[0] == "Kernel Name":
            if take:
                done = True
            take = (kern in r[1]) and not done
            i = 0
            continue
        if r[0] == "Address" or not take:
            continue
        if i < len(lines):
            ex[lines[i]] += int(r[5]); st[lines[i]] += int(r[2])
        i += 1
    tot, stt = sum(ex.values()), sum(st.values())
    src = {}
    print("kernel %s: %d warp instructions, %d stall samples (%d SASS lines mapped)" % (kern, tot, stt, len(lines)))
    for (f, ln), n in sorted(ex.items(), key=lambda kv: -kv[1])[:40]:
        if f not in src:
            p = os.path.join(os.path.dirname(obj), f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][ln - 1].strip()[:110] if 0 < ln <= len(src[f]) else ""
        print("%5.1f%% exec %5.1f%% stall  %s:%d  %s" % (100.0 * n / tot, 100.0 * st[(f, ln)] / max(1, stt), f, ln, text))


if __name__ == "__main__":
    main()
