"""Mirror of the reference's `utils` package for the per-crop hot path.

Module, function and argument names follow /root/reference/utils/{dinov2_utils,feature_util,
projector_util,knn_util,template_util,corresp_util,repre_util}.py so that the per-instance block of
scripts/infer.py (reference :467-545) runs unchanged on top of the sm_100a kernels.
"""
