"""CPU: host-side logic -- C-ABI surface, checkpoint reader, feature builder, trajectory sharding (gloo, world size 2)."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dfmdock_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dfmdock_b200.h")).read()
    declared = set(re.findall(r"\b(dfm_[a-z_0-9]+)\s*\(", header)) - {"dfm_ctx"}
    assert len(declared) >= 18
    lib = _lib.load()                      # symbols only; no GPU needed
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert b"sm_100a" in lib.dfm_version()


def test_header_is_plain_c_and_links(tmp_path):
    """The drop-in boundary is a C ABI: include/dfmdock_b200.h compiles as C99 and as C++11 with -Wall -Wextra -pedantic, and a
    plain-C consumer links against the library and calls entry points that need no GPU (no torch types anywhere)."""
    import shutil
    gcc, gxx = shutil.which("gcc"), shutil.which("g++")
    if not gcc or not gxx:
        pytest.skip("no host compiler")
    from dfmdock_b200 import _lib
    libdir = os.path.dirname(_lib.LIB_PATH)
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "consumer.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "dfmdock_b200.h"\n'
                   'int main(void) {\n'
                   '  if (!dfm_version() || !strstr(dfm_version(), "sm_100a")) return 1;\n'
                   '  if (dfm_set_weight(NULL, "x", NULL, NULL, 0) == 0) return 2;      /* a null context is an error, not a crash */\n'
                   '  if (!dfm_last_error() || !dfm_last_error()[0]) return 3;\n'
                   '  if (dfm_metrics_workspace_bytes(100, 90) == 0) return 4;\n'
                   '  printf("%s\\n", dfm_version());\n  return 0;\n}\n')
    cpp = tmp_path / "consumer.cpp"
    cpp.write_text('#include "dfmdock_b200.h"\nint main() { return dfm_version() ? 0 : 1; }\n')
    flags = ["-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc]
    subprocess.run([gcc, "-std=c99"] + flags + ["-fsyntax-only", str(src)], check=True)
    subprocess.run([gxx, "-std=c++11"] + flags + ["-fsyntax-only", str(cpp)], check=True)
    exe = tmp_path / "consumer"
    subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-ldfmdock_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "dfmdock_b200" in out.stdout


def test_no_cpu_fallback_and_product_never_imports_oracle():
    from dfmdock_b200 import Score_Model
    from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
    m = Score_Model(synthetic_state_dict(0), synthetic_hparams())
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.to("cpu")
    with pytest.raises(RuntimeError):
        m.score(torch.zeros(1, 3, 3, 3), torch.zeros(1))
    pkg = os.path.join(ROOT, "dfmdock_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_lightning_checkpoint_roundtrip(tmp_path):
    from dfmdock_b200.checkpoint import load_checkpoint
    from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict, write_lightning_ckpt
    sd, hp = synthetic_state_dict(3, 67), synthetic_hparams(67)
    p = str(tmp_path / "model.ckpt")
    write_lightning_ckpt(p, sd, hp)
    sd2, hp2 = load_checkpoint(p)
    assert set(sd2) == set(sd) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert hp2["model"]["positional_embed_dim"] == 67 and hp2["diffuser"]["so3"]["max_sigma"] == 1.5
    with pytest.raises(KeyError):
        torch.save({"weights": 1}, p)
        load_checkpoint(p)


@pytest.mark.skipif(not os.path.exists("/root/reference/weights/pinder_0.ckpt"), reason="reference tree not mounted")
def test_shipped_checkpoints_load_unchanged_without_omegaconf():
    from dfmdock_b200.checkpoint import load_checkpoint, load_db5_record
    for path, width in (("/root/reference/weights/pinder_0.ckpt", 67), ("/root/reference/checkpoints/dips/model_0.ckpt", 66)):
        sd, hp = load_checkpoint(path)
        assert len(sd) == 104 and sd["positional_embed.weight"].shape == (128, width)
        assert hp["model"]["node_dim"] == 256 and hp["model"]["cut_off"] == 20.0 and hp["diffuser"]["r3"]["max_sigma"] == 30.0
    assert "omegaconf" not in sys.modules or getattr(sys.modules["omegaconf"], "__file__", None) is None
    rec = load_db5_record("/root/reference/data/db5_test/1QA9.pt")
    assert rec["receptor"]["x"].shape == (102, 1280) and rec["ligand"]["pos"].shape == (95, 3, 3)


def test_features_match_oracle_relpos_and_onehot():
    from dfmdock_b200.features import get_position_matrix, relpos_bins, sequence_to_onehot, synthetic_complex
    from oracle import dfmdock_oracle as orc
    assert torch.equal(relpos_bins(7, 5), orc.relpos_bins(7, 5))
    assert torch.equal(get_position_matrix(40, 30, 67, 1.0), orc.position_matrix(40, 30, 67, 1.0))
    oh = sequence_to_onehot("ACDXZ")
    assert oh.shape == (5, 21) and oh[0, 0] == 1 and oh[3, 20] == 1 and oh[4, 20] == 1
    a, b = synthetic_complex(10, 8, seed=1), synthetic_complex(10, 8, seed=1)
    assert all(torch.equal(a[k], b[k]) for k in a)
    ca = a["rec_pos"][:, 1]
    assert torch.allclose((ca[1:] - ca[:-1]).norm(dim=-1), torch.full((9,), 3.8), atol=1e-4)


def test_diffusers_match_oracle_schedules():
    from dfmdock_b200.diffusers import R3Diffuser, SO3Diffuser
    from oracle import dfmdock_oracle as orc
    so3 = SO3Diffuser({"min_sigma": 0.1, "max_sigma": 1.5, "schedule": "logarithmic"})
    r3 = R3Diffuser({"min_sigma": 0.1, "max_sigma": 30.0})
    for t in (1.0, 0.5, 1e-3):
        assert so3.diffusion_coef(t) == orc.so3_g(t) and r3.diffusion_coef(t) == orc.r3_g(t)
    torch.manual_seed(0)
    a = so3.torch_reverse(torch.ones(1, 3), torch.tensor(0.025), 0.5, noise_scale=0.5)
    torch.manual_seed(0)
    z = 0.5 * torch.randn(1, 3)
    b = orc.reverse_increment(orc.so3_g(0.5), torch.ones(1, 3), torch.tensor(0.025), z)
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        so3.torch_reverse(torch.ones(1, 3), torch.tensor(0.1), torch.tensor([0.5]))


def test_shard_range_is_a_balanced_partition():
    from dfmdock_b200.distributed import shard_range
    for total in (0, 1, 5, 40, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DFM_ROOT"])
from dfmdock_b200.distributed import gather_rows, shard_range, rank_world
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["DFM_PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank, world = rank_world()
total = 7                                     # ragged: 4 + 3
lo, hi = shard_range(total, rank, world)
full_truth = torch.arange(total * 8, dtype=torch.float32).view(total, 8)
out = gather_rows(full_truth[lo:hi].clone(), total)
assert torch.equal(out, full_truth), (rank, out)
best = int(torch.argmin(out[:, 6]))
assert best == 0
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gather_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), DFM_ROOT=ROOT, DFM_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out
        assert "ok" in out


def test_bench_roofline_arithmetic():
    sys.path.insert(0, ROOT)
    import bench
    # 11.66 / 17.76 / 47.36 GFLOP per pose-step at N = 197 / 300 / 800
    for n, gf in ((197, 11.66), (300, 17.76), (800, 47.36)):
        assert abs(bench.algorithmic_flops_per_pose_step(n) / 1e9 - gf) < 0.02
    assert bench.edge_kernel_flops_per_launch(256, 300) == 2 * 4608000 * 65536 + 2 * 4608000 * 256


def test_inference_cli_surface_matches_reference_flags(tmp_path):
    """src/inference.py:569-589: same flag names / defaults; raw PDB inputs without ESM-2 weights are refused loudly."""
    from dfmdock_b200 import inference as inf
    p = inf.build_parser()
    a = p.parse_args(["--paths", "x", "a.pt", "b.pt"])
    assert (a.num_samples, a.num_steps, a.tr_noise_scale, a.rot_noise_scale, a.seed) == (1, 40, 0.5, 0.5, 42)
    assert a.ckpt == "../checkpoints/dips/model_0.ckpt" and a.out_dir == "./pdbs" and a.out_csv == "./test.csv"
    assert not a.use_clash_force and not a.noise_annealing
    with pytest.raises(SystemExit):
        p.parse_args([])                       # --paths | --csv is required, like the reference
    with pytest.raises(RuntimeError, match="esm_dir"):
        inf.load_inputs("1A2K_r_b.pdb", "1A2K_l_b.pdb")
    with pytest.raises(ValueError):
        inf.load_inputs("1A2K_r_b.pdb", "1A2K_l_b.pt")
    assert (a.ode, a.out_trj_dir, a.native_dir, a.esm_dir, a.get_gt_energy) == (False, None, None, None, False)
    # record loading + backbone writer
    from dfmdock_b200.features import synthetic_complex
    b = synthetic_complex(6, 5, seed=1)
    rec = {"receptor": {"x": b["rec_x"][:, :1280], "pos": b["rec_pos"], "seq": "ACDEFG"},
           "ligand": {"x": b["lig_x"][:, :1280], "pos": b["lig_pos"], "seq": "HIKLM"}}
    path = tmp_path / "cplx.pt"
    torch.save(rec, path)
    got = inf.load_inputs(str(path), id="cplx")
    assert got["id"] == "cplx" and got["ligand"]["seq"] == "HIKLM"
    out = tmp_path / "o.pdb"
    inf.write_backbone_pdb(str(out), b["rec_pos"], b["lig_pos"], "ACDEFG", "HIKLM")
    lines = out.read_text().splitlines()
    assert sum(l.startswith("ATOM") for l in lines) == 33 and lines[-1] == "END"
    assert lines[0][12:16].strip() == "N" and lines[0][17:20] == "ALA" and lines[0][21] == "A"
    assert abs(float(lines[1][30:38]) - float(b["rec_pos"][0, 1, 0])) < 1e-3
    assert inf.ligand_rmsd(b["lig_pos"], b["lig_pos"]) == 0.0


def test_plan_work_covers_every_trajectory_once_and_balances():
    """BASELINE config #5 planner: exact cover, deterministic, balanced, never more chunks than useful; sizes as N or as (R, L)."""
    from dfmdock_b200.distributed import complex_cost, plan_work
    db5 = [395, 695, 343, 456, 575, 626, 561, 2548, 329, 197, 352, 320, 430, 377, 430, 382, 339, 628, 404, 240, 535, 373, 588, 492, 214]
    db5_rl = [(223, 172), (368, 327), (242, 101), (311, 145), (470, 105), (426, 200), (432, 129), (2000, 548), (238, 91), (102, 95),
              (223, 129), (195, 125), (269, 161), (170, 207), (355, 75), (275, 107), (275, 64), (574, 54), (263, 141), (120, 120),
              (420, 115), (127, 246), (117, 471), (427, 65), (87, 127)]
    assert [r + l for r, l in db5_rl] == db5
    for sizes, T in ((db5, 40), (db5_rl, 40), ([197], 40), ([300, 300, 300], 7), ([2548], 3), ([], 40)):
        for world in (1, 2, 3, 8):
            plan = plan_work(sizes, T, world)
            assert plan == plan_work(sizes, T, world)
            seen = {}
            load = [0.0] * world
            for c, lo, hi, r in plan:
                assert 0 <= r < world and 0 <= lo < hi <= T
                for k in range(lo, hi):
                    assert (c, k) not in seen
                    seen[(c, k)] = r
                load[r] += complex_cost(sizes[c], hi - lo)
            assert len(seen) == len(sizes) * T
            for c, size in enumerate(sizes):
                n = sum(size) if isinstance(size, tuple) else size
                parts = [ch for ch in plan if ch[0] == c]
                assert len(parts) <= max(1, min(world, T))
                if len(parts) > 1:        # a complex is only cut when the pieces still fill a GPU
                    assert min(hi - lo for _, lo, hi, _ in parts) * n >= 8192 // 2
            if sizes in (db5, db5_rl) and world > 1:
                assert max(load) <= 1.05 * sum(load) / world, (world, load)
    # small complexes stay whole: 40 trajectories x 197 residues is one launch-sized chunk
    assert len(plan_work([197, 214, 240], 40, 8)) == 3
    # the ligand rows cost more than the receptor rows (last layer + coordinate head), the fixed part is paid per chunk
    assert complex_cost((117, 471), 40) > complex_cost((471, 117), 40)
    assert complex_cost(300, 10) + complex_cost(300, 30) > complex_cost(300, 40)


def test_plan_work_against_measured_chunk_times():
    """The planner's cost model against the device times measured on B200 for every db5 complex at 40 / 20 / 10 / 5 trajectories
    (profiles/r02/c5_chunk_times.txt): the model is within 10 % of every measurement, and the 8-rank plan's busiest rank, priced
    with the MEASURED times, stays within 6 % of a perfect split."""
    import re
    from dfmdock_b200.distributed import complex_cost, plan_work
    rl = {"1AVX": (223, 172), "1H1V": (368, 327), "1HCF": (242, 101), "1IRA": (311, 145), "1JIW": (470, 105), "1JPS": (426, 200),
          "1MLC": (432, 129), "1N2C": (2000, 548), "1NW9": (238, 91), "1QA9": (102, 95), "1VFB": (223, 129), "1ZHI": (195, 125),
          "2A1A": (269, 161), "2A9K": (170, 207), "2AYO": (355, 75), "2SIC": (275, 107), "2SNI": (275, 64), "2VDB": (574, 54),
          "3SZK": (263, 141), "4POU": (120, 120), "5C7X": (420, 115), "5HGG": (127, 246), "5JMO": (117, 471), "6B0S": (427, 65),
          "7CEI": (87, 127)}
    ids, times = [], {}
    with open(os.path.join(ROOT, "profiles", "r02", "c5_chunk_times.txt")) as f:
        for line in f:
            m = re.match(r"CHUNK (\S+) N=(\d+)", line)
            if m:
                ids.append(m.group(1))
                assert sum(rl[m.group(1)]) == int(m.group(2))
                times[m.group(1)] = {int(t): float(ms) for t, ms in re.findall(r"T=(\d+) ([\d.]+) ms", line)}
    assert len(ids) == 25
    ms_per_unit = 2.924e-3                       # the fit: time = ms_per_unit * complex_cost (9.8 ms fixed + 2.92 us per row-equivalent)
    for cid in ids:
        for t, ms in times[cid].items():
            assert abs(ms_per_unit * complex_cost(rl[cid], t) - ms) <= 0.10 * ms, (cid, t, ms)

    def measured(cid, n):                        # affine in the trajectory count between / beyond the measured points
        pts = sorted(times[cid].items())
        (t0, m0), (t1, m1) = (pts[0], pts[1]) if n <= pts[1][0] else next((a, b) for a, b in zip(pts, pts[1:]) if n <= b[0])
        return m0 + (m1 - m0) * (n - t0) / (t1 - t0)

    sizes = [rl[c] for c in ids]
    total = sum(times[c][40] for c in ids)
    for world, bound in ((2, 1.03), (4, 1.04), (8, 1.06)):
        load = [0.0] * world
        for c, lo, hi, r in plan_work(sizes, 40, world):
            load[r] += measured(ids[c], hi - lo)
        assert max(load) <= bound * total / world, (world, [round(x) for x in load], total / world)


_GLOO_SET_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DFM_ROOT"])
from dfmdock_b200 import sampler
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["DFM_PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()

class FakeModel:                        # stands in for the CUDA model: result rows encode (complex, trajectory)
    def set_complex(self, batch): self.c = batch["c"]
    def sample(self, lig_pos0, n, stream_base=0, **kw):
        k = torch.arange(stream_base, stream_base + n, dtype=torch.float32)
        return {"lig_pos": (self.c * 1000 + k)[:, None, None, None].expand(n, 2, 3, 3).clone(), "rot_update": k[:, None].expand(n, 3).clone(),
                "tr_update": k[:, None].expand(n, 3).clone(), "energy": -((k - 3 - self.c) ** 2), "num_clashes": torch.zeros(n, dtype=torch.int32)}

loaders = [(lambda c=c: {"c": c, "lig_pos": torch.zeros(2, 3, 3)}) for c in range(4)]
# residue counts -> rows gathered as objects; (R, L) pairs -> one packed float32 all-gather; one complex only -> a rank without work
for sizes in ([900, 200, 2500, 300], [(898, 2), (198, 2), (2498, 2), (298, 2)], [(198, 2)]):
    results, plan = sampler.sample_complex_set(FakeModel(), loaders[:len(sizes)], sizes, 12, num_steps=2, min_nodes=2048)
    assert len(sizes) == 1 or {r for _, _, _, r in plan} == {0, 1}, plan
    assert len(results) == len(sizes)
    for c, res in enumerate(results):
        assert res["lig_pos"].shape == (12, 2, 3, 3)
        assert torch.equal(res["lig_pos"][:, 0, 0, 0], c * 1000 + torch.arange(12, dtype=torch.float32)), (c, res["lig_pos"][:, 0, 0, 0])
        assert torch.equal(res["rot_update"][:, 2], torch.arange(12, dtype=torch.float32))
        assert res["num_clashes"].dtype == torch.int32 and res["energy"].shape == (12,)
        assert res["best"] == int(torch.argmin(res["energy"]))
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_complex_set_sharding_over_gloo(tmp_path):
    script = tmp_path / "worker_set.py"
    script.write_text(_GLOO_SET_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), DFM_ROOT=ROOT, DFM_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
        assert "ok" in out


_GLOO_FEW_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DFM_ROOT"])
from dfmdock_b200 import sampler
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["DFM_PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()

class FakeModel:                        # stands in for the CUDA model: result rows encode the trajectory index
    device = torch.device("cpu")
    def set_complex(self, batch): pass
    def sample(self, lig_pos0, n, stream_base=0, **kw):
        k = torch.arange(stream_base, stream_base + n, dtype=torch.float32)
        return {"lig_pos": k[:, None, None, None].expand(n, 5, 3, 3).clone(), "rot_update": k[:, None].expand(n, 3).clone(),
                "tr_update": k[:, None].expand(n, 3).clone(), "energy": -k, "num_clashes": torch.zeros(n, dtype=torch.int32)}

batch = {"lig_pos": torch.zeros(5, 3, 3)}
# the CLI default: --num_samples 1 under a 2-rank torchrun -> rank 1's shard is EMPTY; both collectives must still complete
for total in (1, 3):
    out = sampler.sample_trajectories(FakeModel(), batch, total, num_steps=2, gather_poses=True)
    assert out["lig_pos"].shape == (total, 5, 3, 3), out["lig_pos"].shape
    assert torch.equal(out["lig_pos"][:, 0, 0, 0], torch.arange(total, dtype=torch.float32))
    assert out["energy"].shape == (total,) and out["best"] == total - 1
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_fewer_trajectories_than_ranks_over_gloo(tmp_path):
    """sample_trajectories(gather_poses=True) with an empty local shard (num_samples < world size)."""
    script = tmp_path / "worker_few.py"
    script.write_text(_GLOO_FEW_WORKER)
    port = str(33500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), DFM_ROOT=ROOT, DFM_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
        assert "ok" in out


def test_unsupported_hparams_are_refused_and_complex_seeds_differ():
    from dfmdock_b200 import Score_Model
    from dfmdock_b200.inference import complex_seed
    from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
    sd = synthetic_state_dict(0)
    for key, bad in (("normalize", False), ("depth", 7), ("node_dim", 128), ("inner_dim", 64)):
        hp = synthetic_hparams()
        hp["model"][key] = bad
        with pytest.raises(ValueError, match=key):
            Score_Model(sd, hp)
    hp = synthetic_hparams()
    hp["diffuser"]["so3"]["schedule"] = "linear"
    with pytest.raises(ValueError):
        Score_Model(sd, hp)
    # trajectory k of different complexes must not share a Philox key; the same (seed, id) must always give the same key
    seeds = {complex_seed(42, cid) for cid in ("1QA9", "7CEI", "4POU", "1N2C")}
    assert len(seeds) == 4 and complex_seed(42, "1QA9") == complex_seed(42, "1QA9") != complex_seed(43, "1QA9")
    assert all(0 <= s < 2 ** 64 for s in seeds)
