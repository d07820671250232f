"""Recipe for baseline/_ref: an UNMODIFIED copy of the parts of the reference checkout that the reference arm of
bench.py executes (`--impl reference`: the reference's own PSFNet.render on the host CPU, or on the GPU as the "eager
before" number) and that BASELINE config 5 needs as the downstream consumer (dff/AiFNet.py).

    python baseline/make_ref.py          # run in the build container, the only place /root/reference exists

The reference has no setup.py / pyproject.toml, so there is nothing to `pip install --target`; its packages are plain
directories and are copied as they lie.  baseline/_ref/ is git-ignored (no reference source enters the history) but
not gpurun-ignored, so it travels to the GPU box with the snapshot.  __graft_entry__.build() runs this when
/root/reference is present.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AADFF_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
ITEMS = ["deeplens", "dff", "lenses/rf50mm", "ckpt/rf50mm/PSFNet480x640_ks11.pkl", "configs",
         "0_warm_up.py", "2_aber_aware_dff_aif.py"]


def main() -> int:
    if not os.path.isdir(REF):
        print(f"{REF} not present; baseline/_ref left as it is ({'exists' if os.path.isdir(DST) else 'missing'})")
        return 0
    for item in ITEMS:
        src, dst = os.path.join(REF, item), os.path.join(DST, item)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copyfile(src, dst)
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write(f"verbatim copy of {ITEMS} from {REF} (singer-yang/Aberration-Aware-Depth-from-Focus), made by baseline/make_ref.py\n")
    print("baseline/_ref written")
    return 0


if __name__ == "__main__":
    sys.exit(main())
