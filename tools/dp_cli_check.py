"""End-to-end check of data-parallel training through the CLI (run on a box with >= 2 GPUs):
    python tools/dp_cli_check.py
makes a CMS workspace with a synthetic table, trains it once with one process and once under torchrun with two ranks
(same global batch), and compares the loss curves and the files written."""
import os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baler_b200 import synth
from baler_b200.modules import helper

CONFIG = '''
def set_config(c):
    c.input_path = "workspaces/CMS_workspace/data/example_CMS_data.npz"
    c.data_dimension = 1
    c.compression_ratio = 1.6
    c.apply_normalization = True
    c.model_name = "%s"
    c.epochs = 3
    c.lr = 0.001
    c.batch_size = 512
    c.early_stopping = True
    c.lr_scheduler = True
    c.early_stopping_patience = 100
    c.min_delta = 0
    c.lr_scheduler_patience = 50
    c.custom_norm = False
    c.reg_param = 0.001
    c.RHO = 0.05
    c.test_size = 0
    c.extra_compression = False
    c.intermittent_model_saving = False
    c.intermittent_saving_patience = 100
    c.activation_extraction = False
    c.deterministic_algorithm = True
    c.separate_model_saving = False
    c.l1 = True
    c.mse_avg = False
    c.mse_sum = True
    c.emd = False
    c.save_error_bounded_deltas = %s
    c.error_bounded_requirement = 25
'''

def main():
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), BALER_B200_SEED="7")
    for model in ("AE", "AE_Dropout_BN"):
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            losses = {}
            for proj, launch in (("single", [sys.executable, "-m", "baler_b200"]),
                                 ("dp2", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                          "--master-addr", "127.0.0.1", "--master-port", "29531", "-m", "baler_b200"])):
                helper.create_new_project("CMS_workspace", proj)
                np.savez("workspaces/CMS_workspace/data/example_CMS_data.npz", data=synth.cms_table(40_000, seed=3), names=synth.CMS_NAMES)
                with open("workspaces/CMS_workspace/%s/config/%s_config.py" % (proj, proj), "w") as f:
                    f.write(CONFIG % (model, "False"))
                r = subprocess.run(launch + ["--project", "CMS_workspace", proj, "--mode", "train"], env=env, capture_output=True, text=True)
                assert r.returncode == 0, "\n".join(l for l in (r.stdout + r.stderr).splitlines() if not l.startswith(('W1', 'I1', '***', 'Setting OMP')))[:6000]
                out = "workspaces/CMS_workspace/%s/output" % proj
                losses[proj] = np.load(out + "/training/loss_data.npy")
                assert os.path.exists(out + "/compressed_output/model.pt") and os.path.exists(out + "/training/normalization_features.npy")
                if model == "AE":  # then compress + decompress with the same launcher, on the model the single-process run trained
                    if proj == "dp2":
                        import shutil
                        for f in ("compressed_output/model.pt", "training/normalization_features.npy"):
                            shutil.copy("workspaces/CMS_workspace/single/output/" + f, out + "/" + f)
                    for mode in ("compress", "decompress"):
                        r = subprocess.run(launch + ["--project", "CMS_workspace", proj, "--mode", mode], env=env, capture_output=True, text=True)
                        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
            if model == "AE":
                for f, key in (("compressed_output/compressed.npz", "data"), ("compressed_output/compressed.npz", "normalization_features"),
                               ("decompressed_output/decompressed.npz", "data")):
                    s1 = np.load("workspaces/CMS_workspace/single/output/" + f)[key]
                    s2 = np.load("workspaces/CMS_workspace/dp2/output/" + f)[key]
                    assert s1.shape == s2.shape and s1.dtype == s2.dtype and np.array_equal(s1, s2), (f, key)
                print("sharded compress / decompress (2 ranks) == single process, bit for bit")
                # the same with the error-bounded-deltas side channel (helper.py:442-470, baler.py:316-338): every rank scans
                # its rows, rank 0 regroups the hits per batch and writes the two gzip'd files
                import gzip
                launches = {"single": [sys.executable, "-m", "baler_b200"],
                            "dp2": [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                    "--master-addr", "127.0.0.1", "--master-port", "29532", "-m", "baler_b200"]}
                for proj, launch in launches.items():
                    with open("workspaces/CMS_workspace/%s/config/%s_config.py" % (proj, proj), "w") as f:
                        f.write(CONFIG % (model, "True"))
                    for mode in ("compress", "decompress"):
                        r = subprocess.run(launch + ["--project", "CMS_workspace", proj, "--mode", mode], env=env, capture_output=True, text=True)
                        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
                n_hits = 0
                for f in ("compressed_deltas.npz.gz", "compressed_batch_index_metadata.npz.gz"):
                    a1 = np.load(gzip.GzipFile("workspaces/CMS_workspace/single/output/compressed_output/" + f), allow_pickle=True)
                    a2 = np.load(gzip.GzipFile("workspaces/CMS_workspace/dp2/output/compressed_output/" + f), allow_pickle=True)
                    assert len(a1) == len(a2)
                    if f.startswith("compressed_deltas"):
                        for d1, d2 in zip(a1, a2):
                            assert np.array_equal(np.asarray(d1), np.asarray(d2))
                            n_hits += len(d1)
                    else:
                        assert np.array_equal(a1[0], a2[0])
                        for (r1, c1), (r2, c2) in zip(a1[1], a2[1]):
                            assert np.array_equal(r1, r2) and np.array_equal(c1, c2)
                s1 = np.load("workspaces/CMS_workspace/single/output/decompressed_output/decompressed.npz")["data"]
                s2 = np.load("workspaces/CMS_workspace/dp2/output/decompressed_output/decompressed.npz")["data"]
                assert np.array_equal(s1, s2) and n_hits > 0
                print("error-bounded deltas: %d hits, files and corrected reconstruction of the 2-rank run == single process" % n_hits)
            a, b = losses["single"][0], losses["dp2"][0]
            print(model, "single", a, "dp2", b, "rel diff", np.abs(a - b) / a)
            assert np.all(np.isfinite(b)) and b[-1] < b[0]
            # same global batch, summed gradients (AE_Dropout_BN: BatchNorm statistics of the global batch and one dropout stream
            # keyed by the global row, exchanged inside the training kernel): the single-process curve within 1 %
            assert np.all(np.abs(a - b) <= 0.01 * a)
    print("dp cli check ok")

if __name__ == "__main__":
    main()
