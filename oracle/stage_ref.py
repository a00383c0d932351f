"""Stage the UNMODIFIED reference's hot-path files under oracle/_ref/ (git-ignored, but it travels to the GPU box
with the working tree like the built libb2c.so).  TEST / BENCH INFRASTRUCTURE: only tests/, bench.py's CPU arms and
__graft_entry__.build() touch oracle/.

    python oracle/stage_ref.py [--reference /root/reference]

What is staged: src/{channel_simulator, baseline_estimators, utils, dataset_generator}.py, the reference's own
test scripts for this path (test_phase1_transmission.py, test_phase2_ls.py, test_phase2_mmse.py) and its config.
Used for (1) `bench.py --impl reference` / `cpu_baseline` with kind = "reference": the reference's own
simulate_transmission + LSEstimator + MMSEEstimator timed on the GPU box's host cores, and (2) a -m gpu test that
runs the reference's test scripts against the drop-in through the `src.` namespace.  Nothing under oracle/_ref/ is
ever committed, and nothing is edited: files are byte-for-byte copies (a manifest of SHA-256 sums is written)."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["src/channel_simulator.py", "src/baseline_estimators.py", "src/utils.py", "src/dataset_generator.py",
         "test_phase1_transmission.py", "test_phase2_ls.py", "test_phase2_mmse.py", "configs/experiment_config.yaml"]


def stage(reference="/root/reference"):
    if not os.path.isdir(reference):
        return None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(reference, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"reference": reference, "sha256": manifest}, fh, indent=1)
    return DEST


def available():
    return all(os.path.exists(os.path.join(DEST, rel)) for rel in FILES)


if __name__ == "__main__":
    ref = sys.argv[sys.argv.index("--reference") + 1] if "--reference" in sys.argv else "/root/reference"
    out = stage(ref)
    print(out if out else f"reference not found at {ref}: nothing staged")
