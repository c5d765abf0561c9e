"""Make-level drop-in proof: the reference's own `ntLink pair` recipe (ntLink:165,198-199,221-225) with this repository's
bin/indexlr and bin/ntlink_pair.py substituted for btllib's indexlr and the reference's ntlink_pair.py -- the seam the
reference actually has (two executables and the files between them). When the unmodified `ntLink` make script is staged in
oracle/_ref (make -C oracle ref) it is run as is, with `ntlink_path=` pointing at this repository's bin/ and bin/ first on
PATH; otherwise the two recipe lines are executed verbatim under `bash -e -o pipefail` (ntLink:91)."""
import os
import shutil
import subprocess
import sys

import pytest

import util

pytestmark = pytest.mark.gpu
BIN = os.path.join(util.REPO, "bin")
NTLINK_MAKE = os.path.join(util.ORACLE_DIR, "_ref", "ntLink")


@pytest.mark.parametrize("case", ["f3_default", "f4_default"])
def test_ntlink_pair_recipe_with_substituted_executables(tmp_path, case):
    man = util.manifest()[case]
    k, w = man["k"], man["w"]
    tgt = shutil.copy(util.fixture_file(tmp_path, man["target"]), str(tmp_path / "target.fa"))
    reads_plain = util.fixture_file(tmp_path, man["reads"])
    reads = str(tmp_path / "reads.fa.gz")
    subprocess.check_call(f"gzip -c {reads_plain} > {reads}", shell=True)
    env = dict(os.environ, PATH=BIN + os.pathsep + os.environ["PATH"], PYTHONPATH=util.REPO + os.pathsep + os.environ.get("PYTHONPATH", ""))
    prefix = f"target.fa.k{k}.w{w}.z1000"
    if os.path.exists(NTLINK_MAKE) and shutil.which("make"):
        cmd = ["make", "-rRf", NTLINK_MAKE, "pair", "target=target.fa", "reads=reads.fa.gz", f"k={k}", f"w={w}", "t=4", "z=1000",
               "ntlink_pairs_tsv=True", "paf=True", f"ntlink_path={BIN}"]
        r = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
        how = "make"
    else:
        script = f"""set -e -o pipefail
indexlr --long --pos --strand -k {k} -w {w} -t 4 target.fa > target.fa.k{k}.w{w}.tsv
sh -c 'gzip -f -cd reads.fa.gz | \\
indexlr --long --pos --strand --len -k {k} -w {w} -t 4 - | \\
{BIN}/ntlink_pair.py -p {prefix} -n 1 -m target.fa.k{k}.w{w}.tsv -s target.fa  \\
-k {k} -a 1 -z 1000 -f 10 -x 0 --verbose --pairs --paf -'
"""
        r = subprocess.run(["bash", "-c", script], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
        how = "recipe lines"
    assert r.returncode == 0, how + "\n" + r.stdout[-2000:] + r.stderr[-3000:]
    out = lambda s: open(os.path.join(str(tmp_path), prefix + s), "rb").read()      # noqa: E731
    assert open(str(tmp_path / f"target.fa.k{k}.w{w}.tsv"), "rb").read() == util.oracle_indexlr(tgt, k, w)
    assert out(".verbose_mapping.tsv") == util.golden_case(case, "verbose_mapping.tsv")
    assert out(".paf") == util.golden_case(case, "paf")
    assert out(".pairs.tsv") == util.golden_case(case, "pairs.tsv")
    assert util.dot_parts(out(".n1.scaffold.dot")) == util.dot_parts(util.golden_case(case, "scaffold.dot"))
