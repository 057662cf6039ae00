"""First-contact GPU script: correctness against the oracle/goldens + coarse timings.  Scratch tool."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import _lib, synth
from mac_b200.solvers import MAC
from mac_b200.utils.fiedler import seeded_start
from oracle import mac_oracle as orc

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def t(): return time.perf_counter()

def check_small():
    fi, fj, fw = synth.complete_graph(5)
    h = _lib.Handle(5, fi, fj, fw, [], [], [])
    h.set_start(seeded_start(5)[:, 0]); h.set_x(np.zeros(0))
    lam, v, info = h.fiedler()
    print("K5 lambda2", lam, info)
    fixed, cand, n = synth.petersen_split()
    gold = json.load(open(os.path.join(G, "petersen.json")))
    for k in range(0, 6):
        mac = MAC(fixed, cand, n)
        x0 = synth.first_k_init(6, k)
        r, w, u = mac.solve(k, x0, max_iters=100)
        gr = gold["runs"][str(k)]
        print("petersen k", k, "dw", np.abs(w - np.array(gr["w"])).max(), "du", u - gr["u"], "rounded ok", (r == np.array(gr["rounded"])).all(),
              "iters", mac.last_info["iters"], len(gr["hist"]))

def check_g2o(name, k):
    z = np.load(os.path.join(G, f"g2o_{name}.npz")); W = np.load(os.path.join(G, "g2o_fw_w.npz"))
    gold = json.load(open(os.path.join(G, "g2o_fw.json")))[name]["runs"][str(k)]
    from mac_b200.g2o import split_edges
    fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"])
    mac = MAC(fixed, cand, n)
    x0 = W[f"{name}_{k}_xinit"]
    t0 = t(); lam, v = mac.fiedler_pair(x0); dt = t() - t0
    v0 = W[f"{name}_{k}_v0"]
    sgn = np.sign(v @ v0)
    print(name, k, "lambda2", lam, "ref", gold["naive_l2"], "rel", abs(lam - gold["naive_l2"]) / gold["naive_l2"], "dv", np.abs(sgn * v - v0).max(), mac.last_info, f"{dt*1e3:.1f} ms")
    t0 = t(); lam, v = mac.fiedler_pair(x0); dt = t() - t0
    print("   second solve", f"{dt*1e3:.1f} ms", mac.last_info)
    t0 = t(); r, w, u = mac.solve(k, x0, max_iters=20); dt = t() - t0
    fh = mac.last_info["f_hist"]; gh = np.array([h["f"] for h in gold["hist"]])
    nn = min(len(fh), len(gh))
    print("   solve", f"{dt*1e3:.1f} ms", "iters", len(fh), gold["iters"], "max rel df", np.abs(fh[:nn] - gh[:nn]).max() / gh.max(), "du", u - gold["u"],
          "dw", np.abs(w - W[f"{name}_{k}_w"]).max(), "rounded diff", int(np.abs(r - W[f"{name}_{k}_rounded"]).sum()))
    print("   counters", mac._h.counters())

def check_er2000():
    gold = json.load(open(os.path.join(G, "er2000.json"))); Z = np.load(os.path.join(G, "er2000.npz"))
    fixed, cand, n = synth.chain_plus_random(2000, 20000, seed=0, weighted=True)
    mac = MAC(fixed, cand, n)
    x0 = synth.first_k_init(20000, 4000)
    f, g = mac.problem(x0)
    print("er2000 f", f, gold["lambda2_init"], "dg", np.abs(g - Z["g0"]).max() / Z["g0"].max(), mac.last_info)
    t0 = t(); w, u, info = mac.frank_wolfe(4000, x0, 10, 0.0, 0.0); dt = t() - t0
    gh = np.array([h["f"] for h in gold["hist"]])
    print("   fw", f"{dt*1e3:.1f} ms", "df", np.abs(info["f_hist"] - gh).max(), "du", u - gold["u"], "dw", np.abs(w - Z["w"]).max())

def check_H():
    t0 = t(); fixed, cand, n, k, x0 = synth.headline(); print("gen H", f"{t()-t0:.1f}s")
    t0 = t(); mac = MAC(fixed, cand, n); print("create H", f"{t()-t0:.2f}s", mac._h.sizes())
    h = mac._h
    t0 = t(); h.set_x(x0); print("set_x", f"{(t()-t0)*1e3:.2f} ms", h.sizes(), "lnorm", h.lnorm())
    v = np.random.default_rng(1).normal(size=n)
    y = h.spmv(v); L = orc.OracleMAC(fixed, cand, n, fw_fiedler_method="arpack").laplacian(x0)
    print("spmv err", np.abs(y - L @ v).max() / np.abs(y).max())
    for rep in range(2):
        t0 = t(); lam, vv, info = h.fiedler(); dt = t() - t0
        print("H fiedler", lam, info, f"{dt*1e3:.2f} ms", "resid_oracle", orc.residual_l1(L, lam, vv))
    ms, by = h.spmv_bench(500); print("spmv L2-resident", f"{ms*1e3:.2f} us", f"{by/ms/1e6:.0f} GB/s algorithmic")
    ms, by = h.spmv_bench(50, True); print("spmv L2-flushed", f"{ms*1e3:.2f} us", f"{by/ms/1e6:.0f} GB/s algorithmic")
    h.reset_counters()
    t0 = t(); w, u, info = mac.frank_wolfe(k, x0, 10, 0.0, 0.0); dt = t() - t0
    print("H fw 10 iters", f"{dt*1e3:.1f} ms", f"{10/dt:.1f} it/s", "f", info["f_hist"], "u", u, h.counters())
    h.set_profile(True); h.reset_counters()
    w, u, info = mac.frank_wolfe(k, x0, 10, 0.0, 0.0)
    print("profiled", h.counters())
    h.set_profile(False)
    for W_ in (4, 8, 16, 32):
        os.environ["MACB_SPMV_W"] = str(W_)
        m2 = MAC(fixed, cand, n); m2._h.set_x(x0)
        ms, by = m2._h.spmv_bench(500); ms2, _ = m2._h.spmv_bench(30, True)
        t0 = t(); lam, _, info = m2._h.fiedler(want_vector=False); dt = t() - t0
        print("W", W_, f"spmv {ms*1e3:.2f} us resident, {ms2*1e3:.2f} us flushed; fiedler {dt*1e3:.2f} ms", info["steps"])
        m2._h.set_x(np.ones(len(x0)) * 0.2)
        ms, by = m2._h.spmv_bench(500)
        t0 = t(); lam, _, info = m2._h.fiedler(want_vector=False); dt = t() - t0
        print("     dense x: spmv", f"{ms*1e3:.2f} us", "fiedler", f"{dt*1e3:.2f} ms", info["steps"], lam)
        m2.close()
    del os.environ["MACB_SPMV_W"]

if __name__ == "__main__":
    what = sys.argv[1:] or ["small", "g2o", "er", "H"]
    if "small" in what: check_small()
    if "g2o" in what:
        check_g2o("intel", 157); check_g2o("sphere2500", 1225); check_g2o("city10000", 1068)
    if "er" in what: check_er2000()
    if "H" in what: check_H()
