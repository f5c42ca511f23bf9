"""LegoLoamParams: the 21 keys of the reference's config/loam_config.yaml as a ctypes POD
(same names; mirrors `struct LegoLoamParams` in include/lego_loam_b200.h)."""
import ctypes as C


class LegoLoamParams(C.Structure):
    _fields_ = [
        ("num_vertical_scans", C.c_int32),
        ("num_horizontal_scans", C.c_int32),
        ("ground_scan_index", C.c_int32),
        ("vertical_angle_bottom", C.c_float),
        ("vertical_angle_top", C.c_float),
        ("sensor_mount_angle", C.c_float),
        ("scan_period", C.c_float),
        ("segment_valid_point_num", C.c_int32),
        ("segment_valid_line_num", C.c_int32),
        ("segment_theta", C.c_float),
        ("edge_threshold", C.c_float),
        ("surf_threshold", C.c_float),
        ("nearest_feature_search_distance", C.c_float),
        ("enable_loop_closure", C.c_int32),
        ("mapping_frequency_divider", C.c_int32),
        ("surrounding_keyframe_search_radius", C.c_float),
        ("surrounding_keyframe_search_num", C.c_int32),
        ("history_keyframe_search_radius", C.c_float),
        ("history_keyframe_search_num", C.c_int32),
        ("history_keyframe_fitness_score", C.c_float),
        ("global_map_visualization_search_radius", C.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def default_params():
    """Values of LeGO-LOAM/config/loam_config.yaml:5-35 (VLP-16)."""
    return LegoLoamParams(
        num_vertical_scans=16, num_horizontal_scans=1800, ground_scan_index=7,
        vertical_angle_bottom=-15.0, vertical_angle_top=15.0, sensor_mount_angle=0.0, scan_period=0.1,
        segment_valid_point_num=5, segment_valid_line_num=3, segment_theta=60.0,
        edge_threshold=0.1, surf_threshold=0.1, nearest_feature_search_distance=5.0,
        enable_loop_closure=0, mapping_frequency_divider=5,
        surrounding_keyframe_search_radius=50.0, surrounding_keyframe_search_num=50,
        history_keyframe_search_radius=7.0, history_keyframe_search_num=25,
        history_keyframe_fitness_score=0.3, global_map_visualization_search_radius=500.0)


def config_params(name):
    """The sensor configurations of BASELINE.json / SURVEY.md section 8d: 'A' VLP-16 16x1800,
    'B' 32x1800, 'C' 64x2048; 'T' is a tiny 16x450 sensor for fast CPU tests."""
    p = default_params()
    if name == "A":
        pass
    elif name == "B":
        p.num_vertical_scans, p.ground_scan_index = 32, 15
    elif name == "C":
        p.num_vertical_scans, p.num_horizontal_scans, p.ground_scan_index = 64, 2048, 31
        p.vertical_angle_bottom, p.vertical_angle_top = -16.6, 16.6
    elif name == "T":
        p.num_horizontal_scans = 450
    else:
        raise ValueError(name)
    return p


def load_yaml(path):
    """Read a loam_config.yaml (same key tree as the reference: lego_loam/{laser,
    imageProjection,featureAssociation,mapping}/*) into a LegoLoamParams."""
    import yaml
    with open(path) as f:
        tree = yaml.safe_load(f)["lego_loam"]
    p = default_params()
    for section in ("laser", "imageProjection", "featureAssociation", "mapping"):
        for k, v in (tree.get(section) or {}).items():
            if not hasattr(p, k):
                raise KeyError(f"unknown parameter {section}/{k}")
            setattr(p, k, int(v) if isinstance(getattr(p, k), int) else float(v))
    return p
