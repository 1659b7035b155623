#!/usr/bin/env python3
"""BASELINE config #1 golden: the UNMODIFIED reference CLI (oracle/_ref/compressor, built by oracle/build_ref.sh from
/root/reference) run through the reference's own bin/run over the survey's 64 MiB mixed dump (524 289 rows x 128 B,
tools/gen_dump.survey_mixed = BASELINE.md section 3) with configs/P6.json.  Build container only.

  tests/golden/cli_survey_P6_results.csv, cli_survey_P6_results_detail.csv, cli_survey_P6_stdout.txt
  tests/golden/cli_survey.json      sha256 of the dump's bytes (the generator depends on numpy's random stream), row count
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bridge import REF_BIN  # noqa: E402
from tools.gen_dump import survey_mixed  # noqa: E402


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    dump = survey_mixed()
    with tempfile.TemporaryDirectory() as tmp:
        ds, out, cwd = (os.path.join(tmp, d) for d in ("ds", "out", "cwd"))
        for d in (ds, out, cwd):
            os.makedirs(d)
        np.save(os.path.join(ds, "survey_mixed.npy"), dump)
        # the reference's bin/run invokes ./compressor from the current directory (bin/run:12)
        shutil.copy(REF_BIN, os.path.join(cwd, "compressor"))
        r = subprocess.run(["bash", "/root/reference/bin/run", "VPC", ds, out, os.path.join(ROOT, "configs", "P6.json")],
                           capture_output=True, text=True, cwd=cwd, check=True)
        with open(os.path.join(gold, "cli_survey_P6_stdout.txt"), "w") as g:
            g.write(r.stdout)
        for suffix in ("results.csv", "results_detail.csv"):
            shutil.copy(os.path.join(out, "P6_" + suffix), os.path.join(gold, "cli_survey_P6_" + suffix))
        print(r.stdout.strip())
    with open(os.path.join(gold, "cli_survey.json"), "w") as g:
        json.dump({"rows": int(dump.shape[0]), "sha256": hashlib.sha256(dump.tobytes()).hexdigest(),
                   "numpy": np.__version__, "generator": "tools/gen_dump.survey_mixed(524289, 1234)"}, g, indent=1)


if __name__ == "__main__":
    main()
