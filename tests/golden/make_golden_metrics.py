"""Generates tests/golden/metrics_*.npz with the UNMODIFIED reference metric functions (/root/reference, CPU):
linnaeus.utils.metrics.basic.accuracy, chain_accuracy.compute_chain_accuracy_vectorized /
compute_partial_chain_accuracy_vectorized, and the per-batch top-1 / top-3 arithmetic of MetricsTracker._update_phase_batch
(tracker.py:697-735, restated inline because the tracker itself needs a full experiment config), plus torch.softmax + torch.topk
as R/inference/handler.py:196-203 calls them.  Build container only:  python tests/golden/make_golden_metrics.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.support import refload  # noqa: E402
from tests.support.golden_metrics import CASES, make_case  # noqa: E402


def main():
    refload.import_reference()
    from linnaeus.utils.metrics.basic import accuracy
    from linnaeus.utils.metrics.chain_accuracy import compute_chain_accuracy_vectorized, compute_partial_chain_accuracy_vectorized

    for name in CASES:
        keys, batches = make_case(name)
        K = len(keys)
        rec = {}
        s1, s3 = np.zeros(K), np.zeros(K)
        nul_ok, nul_n, non_ok, non_n = np.zeros(K), np.zeros(K), np.zeros(K), np.zeros(K)
        chain = partial = tot = 0.0
        for bi, (outputs, targets) in enumerate(batches):
            ol = [torch.from_numpy(outputs[k]) for k in keys]
            tl = [torch.from_numpy(targets[k]) for k in keys]
            B = ol[0].shape[0]
            for i, k in enumerate(keys):
                C = ol[i].shape[1]
                ks = tuple(x for x in (1, 3, 5) if x <= C)
                rec[f"b{bi}.{k}.acc"] = np.array(accuracy(ol[i], tl[i], topk=ks))
                rec[f"b{bi}.{k}.acc_ignore0"] = np.array(accuracy(ol[i], tl[i], topk=ks, ignore_index=0))
                preds = ol[i].argmax(dim=1)  # tracker.py:702-731
                c1 = (preds == tl[i]).sum().item()
                c3 = c1 if min(3, C) < 3 else (ol[i].topk(3, dim=1)[1] == tl[i].unsqueeze(1)).any(dim=1).sum().item()
                s1[i] += c1
                s3[i] += c3
                is_null = tl[i] == 0  # tracker.py:797-912 (hard labels)
                nul_ok[i] += ((preds == tl[i]) & is_null).sum().item()
                nul_n[i] += is_null.sum().item()
                non_ok[i] += ((preds == tl[i]) & ~is_null).sum().item()
                non_n[i] += (~is_null).sum().item()
                kk = min(5, C)
                p, ix = torch.topk(torch.softmax(ol[i], dim=-1), k=kk)
                rec[f"b{bi}.{k}.topk_idx"], rec[f"b{bi}.{k}.topk_prob"] = ix.numpy(), p.numpy()
            ca = compute_chain_accuracy_vectorized(ol, tl)
            pa = compute_partial_chain_accuracy_vectorized(ol, tl)
            rec[f"b{bi}.chain"], rec[f"b{bi}.partial"] = np.array(ca), np.array(pa)
            chain += ca * B  # tracker.py:640-668
            partial += pa * B
            tot += B
        rec["acc1"], rec["acc3"] = 100.0 * s1 / tot, 100.0 * s3 / tot
        with np.errstate(invalid="ignore", divide="ignore"):  # NaN = the tracker records nothing for that task (count 0)
            rec["null_acc1"] = np.where(nul_n > 0, 100.0 * nul_ok / nul_n, np.nan)
            rec["non_null_acc1"] = np.where(non_n > 0, 100.0 * non_ok / non_n, np.nan)
        rec["chain_accuracy"], rec["partial_chain_accuracy"] = np.array(chain / tot), np.array(partial / tot)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        print(name, "acc1", rec["acc1"], "chain", rec["chain_accuracy"], "partial", rec["partial_chain_accuracy"])


if __name__ == "__main__":
    main()
