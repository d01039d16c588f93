"""Verbose (non-asserting) GPU bring-up report: prints CUDA-vs-golden differences stage by stage."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from util import FWD_CASES, load_golden, rel_err, max_abs
from dfmdock_b200 import Score_Model
from oracle import dfmdock_oracle as orc

def main():
    print(torch.cuda.get_device_name(0))
    for name, mk in FWD_CASES.items():
        sd, hp, batch = mk()
        net = orc.OracleNet(sd)
        item = load_golden(name)[0]
        b = dict(batch); b["t"] = torch.tensor([item["t"]])
        keep = {}
        net.forward(b, edges=item["nbr"].long(), keep=keep)
        N = batch["rec_pos"].shape[0] + batch["lig_pos"].shape[0]
        for precision in ("fp32", "fp16"):
            try:
                model = Score_Model(sd, hp, precision=precision).to("cuda")
                model.set_complex(batch)
                t0 = time.time()
                out = model.score(batch["lig_pos"][None], b["t"], edges=item["nbr"][None].int(), want_energy=True)
                torch.cuda.synchronize()
                print("== %s %s (%.1f ms)" % (name, precision, 1e3 * (time.time() - t0)))
                for k in ("f", "tr_score", "rot_score"):
                    print("   %-10s rel %.3e  maxabs %.3e  (ref max %.3e)" % (k, rel_err(out[k].cpu()[0], item[k].reshape(out[k].shape[1:])), max_abs(out[k].cpu()[0], item[k].reshape(out[k].shape[1:])), float(item[k].abs().max())))
                print("   energy %.6f vs %.6f   clashes %d vs %d" % (float(out["energy"][0]), float(item["energy"]), int(out["num_clashes"][0]), int(item["num_clashes"])))
                h = model.debug_read(1, 0, (N, 256)).cpu()
                print("   h5 rel %.3e" % rel_err(h, keep["h5"]))
                agg = model.debug_read(1, 4, (N, 256)).cpu()
                print("   agg5 rel %.3e" % rel_err(agg, keep["agg5"]))
                A = model.debug_read(1, 5, (N, 256)).cpu()
                if hp["model"]["positional_embed_dim"] == 66:
                    W1 = sd["network.EGNN_5.egcl.edge_mlp.0.weight"]
                    Aexp = keep["h4"] @ W1[:, :256].T + sd["network.EGNN_5.egcl.edge_mlp.0.bias"]
                    print("   A5 (W1s h4 + b1) rel %.3e" % rel_err(A, Aexp))
                ft = model.debug_read(1, 1, (N, 64), dtype=torch.int32).cpu()
                K = item["nbr"].shape[1]
                rows = torch.arange(N)[:, None].expand(N, K)
                gb = item["bins"].long(); nb = item["nbr"].long()
                ftl = ft[:, :K].long()
                for q, (sh, mk_) in enumerate(((0, 63), (6, 31), (11, 31), (16, 15))):
                    x = (ftl >> sh) & mk_
                    print("   bins[%d] mismatches %d / %d" % (q, int((x != gb[q][rows, nb]).sum()), N * K))
            except Exception:
                traceback.print_exc()

if __name__ == "__main__":
    main()
