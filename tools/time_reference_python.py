"""Times the REFERENCE's own Python (imported unmodified from /root/reference) as BASELINE.md section 3 plans:
build_suffix_array at n = 64 K / 128 K / 200 K (csa/suffix_array.py:131-134; O(n^2) bytes of suffix copies -- the wall
that excludes 1 MB and up), and at the C1 size (1 MiB ENG96 text) the other stages with the suffix array injected
from the oracle: bwt_transform, build_count, build_occ, WaveletTree(bwt), then find_range / find over 1 k patterns.
Runs in the authoring container only (needs /root/reference; single-threaded Python, 1 core).  Writes one JSON file.

    PYTHONDONTWRITEBYTECODE=1 python tools/time_reference_python.py profiles/r02_reference_python_timings.json
"""
import contextlib, io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HKCSA_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
with contextlib.redirect_stdout(io.StringIO()):
    from csa.suffix_array import build_suffix_array
    from csa.bwt import bwt_transform
    from csa.wavelet_tree import WaveletTree
    import csa.enhanced_fm_index as efm
    from utils.utils import build_count, build_occ
from oracle import oracle as O


def timed(fn, *a):
    t0 = time.perf_counter()
    r = fn(*a)
    return r, time.perf_counter() - t0


def main(out_path):
    res = {"host": {"cpu_count": os.cpu_count(), "python": sys.version.split()[0]},
           "what": "reference Python, unmodified, 1 core; texts from the oracle's seeded ENG96 generator (seed 42)",
           "build_suffix_array": []}
    for n in (65_536, 131_072, 200_000):
        text = O.gen_text(O.ENG96, 42, n).tobytes().decode("latin-1")
        sa, dt = timed(build_suffix_array, text)
        assert sa == O.build_suffix_array(text.encode("latin-1")).tolist()
        res["build_suffix_array"].append({"n": n, "seconds": dt, "MB_per_s": n / 1e6 / dt,
                                          "suffix_copy_bytes": n * (n + 1) // 2})
        print(res["build_suffix_array"][-1], flush=True)
        del sa
    n = 1 << 20
    raw = O.gen_text(O.ENG96, 42, n).tobytes()
    text = raw.decode("latin-1")
    t_sa0 = time.perf_counter()
    sa_np = O.build_suffix_array(raw + b"$")
    t_oracle_sa = time.perf_counter() - t_sa0
    sa = sa_np.tolist()
    stages = {}
    tx = text + "$"
    bwt, stages["bwt_transform"] = timed(bwt_transform, tx, sa)
    _, stages["build_count"] = timed(build_count, tx)
    occ, stages["build_occ"] = timed(build_occ, bwt)
    del occ
    with contextlib.redirect_stdout(io.StringIO()):
        _, stages["WaveletTree(bwt)"] = timed(WaveletTree, bwt)
    # EnhancedFMIndex with the suffix array injected (its own build_suffix_array cannot run at 1 MiB)
    efm.build_suffix_array = lambda t: sa
    fm, stages["EnhancedFMIndex(text) with injected SA"] = timed(efm.EnhancedFMIndex, text)
    pats, off = O.gen_patterns(44, 1000, raw)
    qs = [pats[off[k]:off[k + 1]].tobytes().decode("latin-1") for k in range(1000)]
    t0 = time.perf_counter()
    ranges = [fm.find_range(q) for q in qs]
    stages["find_range x 1000"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    occs = sum(len(fm.find(q)) for q in qs)
    stages["find x 1000"] = time.perf_counter() - t0
    res["c1_1MiB"] = {"n": n + 1, "stages_seconds": stages, "oracle_suffix_sort_seconds_for_the_injected_SA": t_oracle_sa,
                      "find_range_patterns_per_s": 1000 / stages["find_range x 1000"],
                      "hits": sum(1 for r in ranges if r[0] >= 0), "located_occurrences": occs,
                      "build_without_SA_MB_per_s": n / 1e6 / (stages["bwt_transform"] + stages["build_count"] + stages["build_occ"])}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res["c1_1MiB"]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_reference_python_timings.json")
