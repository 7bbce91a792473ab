"""One MAQ pass (findSNPs over makePileup(reads, use_md_tag)) on a synthetic file, for an ncu capture of the kernels of
rows N1 / N3 (md_len_kernel, md_replay_kernel, maq_kernel):
    ncu --set full -k regex:"maq_kernel|md_len_kernel|md_replay_kernel" -c 6 -o gpurun_out/prof_maq python tools/maq_profile.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    from tools import bamgen
    from biod_b200 import BamReader, MaqSnpCaller
    data = bamgen.generate(n, 1, False, -1, bamgen.SEED_BASE + 2)
    rd = BamReader(data.tobytes())
    t = time.perf_counter()
    calls = sum(1 for _ in MaqSnpCaller().findSNPs(rd))
    print(f"{n} reads, {calls} calls, {time.perf_counter() - t:.2f} s")


if __name__ == "__main__":
    main()
