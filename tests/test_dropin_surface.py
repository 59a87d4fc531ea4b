"""The drop-in boundary: the public methods of `OVO`, `CLIPGenerator`, `MaskGenerator`, `Instance3D` and
`match_labels_to_vtx` take the reference's parameters, by the same names, in the same order, with the same defaults.
Compared against the reference's own source when it is present (the build container); skipped on the GPU box, where
/root/reference does not exist.  The reference modules are parsed with `ast` (importing them needs un-vendored packages)."""
import ast
import inspect
import os

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ovo")), reason="reference tree not available")


def ref_signatures(path, cls=None):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)       # the reference's docstrings hold "\i" escapes
        tree = ast.parse(open(os.path.join(REF, path)).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    out = {}
    for n in body:
        if isinstance(n, ast.FunctionDef):
            a = n.args
            names = [x.arg for x in a.args]
            defaults = [ast.literal_eval(d) if not isinstance(d, (ast.Name, ast.Attribute, ast.Call)) else "<expr>" for d in a.defaults]
            out[n.name] = (names, defaults)
    return out


def our_signature(fn):
    sig = inspect.signature(fn)
    names = [p.name for p in sig.parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    defaults = [p.default for p in sig.parameters.values() if p.default is not p.empty and p.kind == p.POSITIONAL_OR_KEYWORD]
    return names, defaults


def check(ours_cls_or_mod, ref, methods):
    for m in methods:
        r_names, r_defaults = ref[m]
        o_names, o_defaults = our_signature(getattr(ours_cls_or_mod, m))
        # ours may append optional parameters AFTER the reference's (an injected encoder, a state_dict, ...), never before
        assert o_names[: len(r_names)] == r_names, (m, o_names, r_names)
        n_required_ref = len(r_names) - len(r_defaults)
        assert len(o_names) - len(o_defaults) == n_required_ref, (m, "required parameters differ")
        for a, b in zip(r_defaults, o_defaults[: len(r_defaults)]):
            if a != "<expr>":
                assert a == b, (m, a, b)


def test_ovo_surface():
    from ovo_b200.ovo import OVO
    ref = ref_signatures("ovo/entities/ovo.py", "OVO")
    check(OVO, ref, ["__init__", "detect_and_track_objects", "compute_semantic_info", "complete_semantic_info", "update_map", "query",
                     "classify_instances", "get_objs_clips", "capture_dict", "restore_dict", "cpu", "cuda", "to", "update_objects_clip"])


def test_clip_generator_surface():
    from ovo_b200.clip_generator import CLIPGenerator
    ref = ref_signatures("ovo/entities/clip_generator.py", "CLIPGenerator")
    check(CLIPGenerator, ref, ["__init__", "extract_clip", "get_txt_embedding", "get_embed_txt_similarity", "cpu", "cuda", "to"])


def test_mask_generator_surface():
    from ovo_b200.mask_generator import MaskGenerator
    ref = ref_signatures("ovo/entities/mask_generator.py", "MaskGenerator")
    check(MaskGenerator, ref, ["__init__", "get_masks", "segment", "precompute", "cpu", "cuda", "to"])


def test_instance3d_surface():
    from ovo_b200.instance3d import Instance3D
    ref = ref_signatures("ovo/entities/instance3d.py", "Instance3D")
    check(Instance3D, ref, ["__init__", "update", "add_points_ids", "add_keyframes", "add_top_kf", "is_top_kf", "idx_in_top_kf",
                            "update_clip", "export", "restore", "purge_points_ids"])


def test_eval_utils_and_mapper_surface():
    from ovo_b200 import eval_utils
    from ovo_b200.mapper import PointMapper
    check(eval_utils, ref_signatures("ovo/utils/eval_utils.py"), ["match_labels_to_vtx"])
    ref = ref_signatures("ovo/slam/vanilla_mapper.py", "VanillaMapper")
    check(PointMapper, ref, ["__init__", "track_camera", "map", "get_c2w", "get_map", "get_kfs", "update_pcd_obj_ids", "get_pcd_colors",
                             "get_map_dict", "set_map_dict", "get_cam_dict", "set_cam_dict"])
