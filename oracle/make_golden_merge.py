"""Golden output of the reference's OWN ``merge_pdbfiles`` (slm/utils/eval_utils.py:437-492, exec'd from its source)
on single-model inputs -- what sample_esmdiff.py:231 produces from the per-sample files.  Pins the framing of the
multi-MODEL file (MODEL n / ATOM / TER / ENDMDL per input, then a closing ENDMDL and END, 80-column lines).
TEST INFRASTRUCTURE; run in the build container:   python -m oracle.make_golden_merge  ->  tests/golden/merged_models.pdb"""
import tempfile
from pathlib import Path

import numpy as np

from . import ref_loader


def inputs():
    """Three single-model PDB texts as ``ESMProtein.to_pdb`` leaves them (ATOM records, TER, END)."""
    from esmdiff_b200.decoder import pdb_model_lines
    rng = np.random.default_rng(3)
    out = []
    for n in range(3):
        bb = (rng.standard_normal((5, 3, 3)) * 9).astype(np.float32)
        o = (rng.standard_normal((5, 3)) * 9).astype(np.float32)
        o[-1] = np.nan
        lines = pdb_model_lines("ACD_K", bb, o, rng.random(5).astype(np.float32))
        out.append(("ACD_K", bb, o, "\n".join(lines + ["END"]) + "\n"))
    return out


def main():
    src = ref_loader._extract(ref_loader.REF_ROOT / "slm/utils/eval_utils.py", {"merge_pdbfiles"})
    ns = {"Path": Path, "tqdm": lambda it, **k: it}
    exec(src, ns)
    with tempfile.TemporaryDirectory() as tmp:
        files = []
        for i, (_, _, _, text) in enumerate(inputs()):
            f = Path(tmp) / f"s.{i}.pdb"
            f.write_text(text)
            files.append(f)
        out = Path(tmp) / "merged.pdb"
        ns["merge_pdbfiles"](files, out, verbose=False)
        gold = Path(__file__).resolve().parent.parent / "tests" / "golden" / "merged_models.pdb"
        gold.write_text(out.read_text())
    print(gold, gold.read_text().splitlines()[-3:])


if __name__ == "__main__":
    main()
