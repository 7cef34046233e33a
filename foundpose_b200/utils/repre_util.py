"""Object representation ("bank") container and repre.pth I/O - mirror of the reference's utils/repre_util.py.

Same dataclass fields, same `repre.pth` dictionary layout (torch.save of tensors + option dicts +
projector tensordicts + camera dicts, reference :99-141), so files written by the reference's
gen_repre.py load here and vice versa.  Template cameras are kept as plain dicts
({"f", "c", "width", "height", "T_world_from_eye"}): the hot path never reads them
and the reference's camera classes (utils/structs.py) are out of scope.
"""

import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, NamedTuple, Optional

import torch

from foundpose_b200.utils import logging, projector_util, structs
from foundpose_b200.utils.misc import tensor_to_array

logger: logging.Logger = logging.get_logger()


class FeatureOpts(NamedTuple):
    extractor_name: str


class TemplateDescOpts(NamedTuple):
    desc_type: str = "tfidf"

    # Options for tfidf template descriptor.
    tfidf_knn_metric: str = "l2"
    tfidf_knn_k: int = 3
    tfidf_soft_assign: bool = False
    tfidf_soft_sigma_squared: float = 10.0


@dataclass
class FeatureBasedObjectRepre:
    """Stores visual object features registered in 3D."""

    vertices: Optional[torch.Tensor] = None               # (num_vertices, 3)
    vertex_normals: Optional[torch.Tensor] = None         # (num_vertices, 3)
    feat_vectors: Optional[torch.Tensor] = None           # (num_features, feat_dims)
    feat_opts: Optional[FeatureOpts] = None
    feat_to_vertex_ids: Optional[torch.Tensor] = None     # (num_features)
    feat_to_template_ids: Optional[torch.Tensor] = None   # (num_features)
    feat_to_cluster_ids: Optional[torch.Tensor] = None    # (num_features)
    feat_cluster_centroids: Optional[torch.Tensor] = None  # (num_clusters, feat_dims)
    feat_cluster_idfs: Optional[torch.Tensor] = None      # (num_clusters)
    feat_raw_projectors: List[projector_util.Projector] = field(default_factory=list)
    feat_vis_projectors: List[projector_util.Projector] = field(default_factory=list)
    templates: Optional[torch.Tensor] = None              # (num_templates, channels, height, width)
    template_cameras_cam_from_model: List[Any] = field(default_factory=list)
    template_descs: Optional[torch.Tensor] = None         # (num_templates, desc_dims)
    template_desc_opts: Optional[TemplateDescOpts] = None


def get_object_repre_dir_path(base_dir: str, repre_type: str, dataset: str, lid: int) -> str:
    """Path to the directory where a representation of the specified object is stored."""
    return os.path.join(base_dir, dataset, repre_type, str(lid))


def _camera_to_dict(camera: Any) -> Dict[str, Any]:
    if isinstance(camera, dict):
        return camera
    return {
        "f": torch.as_tensor(camera.f),
        "c": torch.as_tensor(camera.c),
        "width": camera.width,
        "height": camera.height,
        "T_world_from_eye": torch.as_tensor(camera.T_world_from_eye),
    }


def save_object_repre(repre: FeatureBasedObjectRepre, repre_dir: str) -> None:
    object_dict: Dict[str, Any] = {}
    for key, value in repre.__dict__.items():
        if key.startswith("_"):
            continue
        if value is not None and torch.is_tensor(value):
            object_dict[key] = value.detach().cpu()
    object_dict["template_cameras_cam_from_model"] = [
        _camera_to_dict(c) for c in repre.template_cameras_cam_from_model
    ]
    object_dict["feat_opts"] = repre.feat_opts._asdict()
    object_dict["template_desc_opts"] = repre.template_desc_opts._asdict()
    object_dict["feat_raw_projectors"] = [
        projector_util.projector_to_tensordict(p) for p in repre.feat_raw_projectors
    ]
    object_dict["feat_vis_projectors"] = [
        projector_util.projector_to_tensordict(p) for p in repre.feat_vis_projectors
    ]
    os.makedirs(repre_dir, exist_ok=True)
    repre_path = os.path.join(repre_dir, "repre.pth")
    logger.info(f"Saving repre to: {repre_path}")
    torch.save(object_dict, repre_path)


def load_object_repre(repre_dir: str, tensor_device: str = "cuda",
                      load_fields: Optional[List[str]] = None) -> FeatureBasedObjectRepre:
    """Loads a representation of the specified object.

    As in the reference (:203-208, device move commented out) tensors stay where torch.load puts
    them (CPU); the packed device layout is built by pipeline.ObjectIndex on first use.
    """
    repre_path = os.path.join(repre_dir, "repre.pth")
    logger.info(f"Loading repre from: {repre_path}")
    object_dict = torch.load(repre_path, map_location="cpu", weights_only=False)
    logger.info("Repre loaded.")

    repre_dict: Dict[str, Any] = {}
    for key, value in object_dict.items():
        if value is not None and isinstance(value, torch.Tensor):
            repre_dict[key] = value
    if object_dict.get("feat_opts") is not None and (load_fields is None or "feat_opts" in load_fields):
        repre_dict["feat_opts"] = FeatureOpts(**dict(object_dict["feat_opts"]))
    repre_dict["feat_raw_projectors"] = []
    if load_fields is None or "feat_raw_projectors" in load_fields:
        for projector in object_dict.get("feat_raw_projectors", []):
            repre_dict["feat_raw_projectors"].append(projector_util.projector_from_tensordict(projector))
    repre_dict["feat_vis_projectors"] = []
    if load_fields is None or "feat_vis_projectors" in load_fields:
        for projector in object_dict.get("feat_vis_projectors", []):
            repre_dict["feat_vis_projectors"].append(projector_util.projector_from_tensordict(projector))
    repre_dict["template_cameras_cam_from_model"] = []
    if load_fields is None or "template_cameras_cam_from_model" in load_fields:
        # Same return structure as the reference (:179-190): PinholePlaneCameraModel objects.
        for camera in object_dict.get("template_cameras_cam_from_model", []):
            repre_dict["template_cameras_cam_from_model"].append(structs.PinholePlaneCameraModel(
                f=tuple(float(v) for v in torch.as_tensor(camera["f"]).flatten().tolist()),
                c=tuple(float(v) for v in torch.as_tensor(camera["c"]).flatten().tolist()),
                width=int(camera["width"]), height=int(camera["height"]),
                T_world_from_eye=torch.as_tensor(camera["T_world_from_eye"]).to(torch.float64).numpy()))
    if load_fields is None or "template_desc_opts" in load_fields:
        if object_dict.get("template_desc_opts") is not None:
            repre_dict["template_desc_opts"] = TemplateDescOpts(**dict(object_dict["template_desc_opts"]))
    return FeatureBasedObjectRepre(**repre_dict)


def convert_object_repre_to_numpy(repre: FeatureBasedObjectRepre) -> FeatureBasedObjectRepre:
    repre_out = FeatureBasedObjectRepre()
    for name, value in repre.__dict__.items():
        if name.startswith("_"):
            continue
        if value is not None and isinstance(value, torch.Tensor):
            value = tensor_to_array(value)
        setattr(repre_out, name, value)
    return repre_out
