"""On-disk traces in the reference's layout, so its own post-processing
(`surrDAMH/modules/visualization_and_analysis.py: Samples.load_MH / load_notes / load_accepted /
load_rejected / plot_raw_data`) can read a GPU run.

Per stage i and chain c the reference's sampler writes (classes_SAMPLER.py:58-69,107-118,130-134,
169-177,196-214), with name = "alg%03d%s_rank%d" % (i, type, c):
  saved_samples/<name>/data/<name>.csv           rows [weight, u_1..u_d, posterior, pre_posterior]: the compressed
                                                  chain -- one row per displaced sample, plus the last one
  saved_samples/<name>/raw_data/<name>.csv        rows ['accepted' | 'rejected' | 'prerejected', u'_1..u'_d]
  saved_samples/<name>/notes/<name>.csv           header + one row of 10 labelled fields
  saved_samples/<name>/DAMH_accepted|DAMH_rejected/<name>.csv   rows [step, posterior', pre_posterior']  (DAMH only)

Use as the `trace_sink` of LockstepSampler.  One set of files per chain: meant for the chain counts the
reference itself is run with (`max_chains` guards against millions of files); large runs should keep the
device-side moments instead.
"""
import csv
import os

import numpy as np

from . import _lib as L

_NAMES = {L.FLAG_PREREJECTED: "prerejected", L.FLAG_ACCEPTED: "accepted", L.FLAG_REJECTED: "rejected"}


class CsvTraceWriter:
    def __init__(self, saved_samples_name, root=".", max_chains=1024):
        self.base = os.path.join(root, "saved_samples", saved_samples_name)
        self.max_chains = max_chains
        self._st = None

    # ---- LockstepSampler hooks ------------------------------------------------------------------
    def begin_stage(self, i, st, sampler):
        C_ = sampler.C
        if C_ > self.max_chains:
            raise L.SdbError(L.ERR_LIMIT, "CsvTraceWriter: %d chains would create %d files; raise max_chains or keep "
                             "device-side statistics" % (C_, 5 * C_))
        blk = sampler.block
        kind = st["type"]
        names = ["alg%03d%s_rank%d" % (i, kind, sampler.chain0 + c) for c in range(C_)]
        dirs = ["data", "raw_data"] + (["DAMH_accepted", "DAMH_rejected"] if kind == "DAMH" else [])
        files = {}
        for dname in dirs:
            os.makedirs(os.path.join(self.base, dname), exist_ok=True)
            files[dname] = [open(os.path.join(self.base, dname, n + ".csv"), "w", newline="") for n in names]
        self._st = dict(i=i, kind=kind, names=names, files=files, writers={k: [csv.writer(f) for f in v] for k, v in files.items()},
                        cur=blk.current_numpy(), post=blk.post_u.cpu().numpy().copy(), pre=blk.pre_u.cpu().numpy().copy(),
                        rej=np.zeros(C_, dtype=np.int64), step=0, st=st, seed=sampler.conf.seed)

    def __call__(self, stage, step0, K, trace):
        s = self._st
        flags = trace.flags[:K].cpu().numpy()
        prop = trace.prop[:K].cpu().numpy()          # [K][d][C]
        post = trace.post[:K].cpu().numpy()
        pre = trace.pre[:K].cpu().numpy()
        pre_cur = trace.pre_cur[:K].cpu().numpy()
        damh = s["kind"] == "DAMH"
        W = s["writers"]
        for k in range(K):
            for c in range(flags.shape[1]):
                f = int(flags[k, c])
                p = prop[k, :, c]
                if damh:
                    s["pre"][c] = pre_cur[k, c]      # recomputed every step by the reference (:189)
                W["raw_data"][c].writerow([_NAMES[f]] + list(p))
                if f == L.FLAG_ACCEPTED:
                    W["data"][c].writerow([1 + s["rej"][c]] + list(s["cur"][c]) + [s["post"][c], s["pre"][c]])
                    if damh:
                        W["DAMH_accepted"][c].writerow([s["step"] + k, post[k, c], pre[k, c]])
                    s["cur"][c], s["post"][c], s["rej"][c] = p.copy(), post[k, c], 0
                else:
                    s["rej"][c] += 1
                    if damh and f == L.FLAG_REJECTED:
                        W["DAMH_rejected"][c].writerow([s["step"] + k, post[k, c], pre[k, c]])
        s["step"] += K

    def end_stage(self, i, st, sampler, steps_done):
        s = self._st
        counters = sampler.block.counters_numpy()
        os.makedirs(os.path.join(self.base, "notes"), exist_ok=True)
        labels = ["name", "surrogate_is_updated", "no_accepted", "no_rejected", "no_prerejected", "no_all", "acceptance_rate",
                  "proposal_std", "proposal_scale", "seed"]
        for c, name in enumerate(s["names"]):
            s["writers"]["data"][c].writerow([1 + s["rej"][c]] + list(s["cur"][c]) + [s["post"][c], s["pre"][c]])   # close_files (:104)
            acc, rej, pre = [int(v) for v in counters[c][:3]]
            no_all = acc + rej + pre
            std = st["proposal_std"][sampler.chain0 + c] if st.get("proposal_differs", False) else st["proposal_std"]
            scale = np.array(std, dtype=float).reshape((-1,))[0]
            with open(os.path.join(self.base, "notes", name + ".csv"), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(labels)
                w.writerow([name, st["surrogate_is_updated"], acc, rej, pre, no_all, acc / max(1, no_all), std, scale, s["seed"]])
        for fl in s["files"].values():
            for f in fl:
                f.close()
        self._st = None
