"""Generates tests/golden/ref_repre/repre.pth with the REFERENCE'S OWN `save_object_repre`, and checks that the
reference's own `load_object_repre` reads a file written by this repo's writer.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_repre.py

The representation is the small bank of tests/golden/make_golden.py (24 templates, 64-d, 32 visual words, seed 51)
- the one `golden_v1.pt`'s correspondences were produced from - plus one PCA projector and two template cameras
(reference `PinholePlaneCameraModel`), written by utils/repre_util.py:99-141 of the reference.  The GPU tests load
that file with foundpose_b200.utils.repre_util.load_object_repre and must reproduce the golden correspondences.
The cross-load result (reference loader on our file) is recorded in ref_repre/cross_load.json.
"""

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402  (sets sys.path for the reference, provides the stubs)

from foundpose_b200 import synthetic  # noqa: E402


def main() -> None:
    make_golden.install_stubs()
    from utils import projector_util, repre_util, structs

    gold = torch.load(os.path.join(HERE, "golden_v1.pt"), weights_only=False)
    bank = synthetic.make_bank_tensors(num_templates=24, patches_per_template=48, feat_dim=64, num_words=32,
                                       seed=gold["bank/seed"], ragged=True)
    rng = np.random.RandomState(7)
    cameras = []
    for i in range(2):
        T = synthetic._rigid_transform(rng)
        cameras.append(structs.PinholePlaneCameraModel(width=420, height=420, f=(600.0 + i, 601.0 + i),
                                                       c=(209.5, 210.5 + i), T_world_from_eye=T))
    projector = projector_util.projector_from_tensordict(synthetic.make_pca(128, 64, seed=41))
    repre = repre_util.FeatureBasedObjectRepre(
        vertices=bank["vertices"], feat_vectors=bank["feat_vectors"],
        feat_to_template_ids=bank["feat_to_template_ids"], feat_to_vertex_ids=bank["feat_to_vertex_ids"],
        feat_to_cluster_ids=gold["bank/feat_to_word"].to(torch.int32),
        feat_cluster_centroids=bank["feat_cluster_centroids"], feat_cluster_idfs=gold["bank/idfs"],
        template_descs=gold["bank/template_descs"], template_desc_opts=repre_util.TemplateDescOpts(),
        feat_opts=repre_util.FeatureOpts(extractor_name="dinov2_vits14-reg"),
        feat_raw_projectors=[projector], template_cameras_cam_from_model=cameras)
    out_dir = os.path.join(HERE, "ref_repre")
    os.makedirs(out_dir, exist_ok=True)
    repre_util.save_object_repre(repre, out_dir)                       # the reference's writer
    size = os.path.getsize(os.path.join(out_dir, "repre.pth"))

    # ---- cross-load: the reference's loader on a file written by this repo -----------------------
    from foundpose_b200.utils import repre_util as ours

    theirs_back = repre_util.load_object_repre(out_dir)                # sanity: reference reads its own file
    ours_repre = ours.load_object_repre(out_dir)                       # our loader on the reference's file
    with tempfile.TemporaryDirectory() as tmp:
        ours.save_object_repre(ours_repre, tmp)                        # our writer
        ref_reads_ours = repre_util.load_object_repre(tmp)             # the reference's loader
    checks = {}
    for name in ("vertices", "feat_vectors", "feat_to_template_ids", "feat_to_vertex_ids", "feat_to_cluster_ids",
                 "feat_cluster_centroids", "feat_cluster_idfs", "template_descs"):
        checks[name] = bool(torch.equal(getattr(ref_reads_ours, name), getattr(theirs_back, name)))
    checks["feat_opts"] = ref_reads_ours.feat_opts == theirs_back.feat_opts
    checks["template_desc_opts"] = ref_reads_ours.template_desc_opts == theirs_back.template_desc_opts
    checks["projector_components"] = bool(np.array_equal(ref_reads_ours.feat_raw_projectors[0].pca.components_,
                                                         theirs_back.feat_raw_projectors[0].pca.components_))
    checks["projector_mean"] = bool(np.array_equal(ref_reads_ours.feat_raw_projectors[0].pca.mean_,
                                                   theirs_back.feat_raw_projectors[0].pca.mean_))
    cam_a, cam_b = ref_reads_ours.template_cameras_cam_from_model, theirs_back.template_cameras_cam_from_model
    checks["cameras"] = len(cam_a) == len(cam_b) == 2 and all(
        type(a).__name__ == "PinholePlaneCameraModel" and np.allclose(np.asarray(a.f, dtype=np.float64),
                                                                      np.asarray(b.f, dtype=np.float64))
        and np.allclose(np.asarray(a.c, dtype=np.float64), np.asarray(b.c, dtype=np.float64))
        and a.width == b.width and a.height == b.height
        and np.allclose(np.asarray(a.T_world_from_eye), np.asarray(b.T_world_from_eye)) for a, b in zip(cam_a, cam_b))
    assert all(checks.values()), checks
    with open(os.path.join(out_dir, "cross_load.json"), "w") as f:
        json.dump({"reference_loader_reads_our_file": checks, "reference_written_bytes": size,
                   "torch": torch.__version__}, f, indent=1)
    print(f"wrote {out_dir}/repre.pth ({size / 1e3:.0f} kB); reference loader on our file: all fields equal")


if __name__ == "__main__":
    main()
