"""The `model:` block of the reference's configs/switch_nerf/building.yaml and the hparams Namespace
fields the hot path reads -- the input spec of the path, as plain Python."""
from argparse import Namespace


def building_model_cfg(width=256, experts_layers=7, skips=(3,), appearance_dim=48,
                       pos_xyz_dim=12, pos_dir_dim=4, hidden2=128):
    """The `model:` block of configs/switch_nerf/building.yaml (lines 6-83),
    parameterised by width so that small plumbing configs share the topology."""
    xyz_in = 3 + 3 * 2 * pos_xyz_dim
    dir_in = 3 + 3 * 2 * pos_dir_dim
    return {
        "layer_num_main": 3, "sigma_tag": 0, "dir_tag": 1, "color_tag": 2,
        "layers": {
            "xyz": {"in_ch": xyz_in, "h_ch": 0, "out_ch": width, "num": 1, "type": "mlp", "act": "none"},
            "0": {"in_ch": width, "h_ch": width, "out_ch": width, "num": experts_layers,
                  "skips": list(skips), "init_factor": 1.0, "type": "moe", "act": "relu",
                  "gate_type": "top", "k": 1, "fp32_gate": True, "gate_dim": width},
            "1": {"in_ch": width, "h_ch": 0, "out_ch": width, "num": 1, "type": "mlp", "act": "none"},
            "2": {"in_ch": width + dir_in + appearance_dim, "h_ch": 0, "out_ch": hidden2, "num": 1,
                  "type": "mlp", "act": "relu"},
            "sigma": {"in_ch": width, "h_ch": 0, "out_ch": 1, "num": 1, "type": "mlp", "act": "none"},
            "color": {"in_ch": hidden2, "h_ch": 0, "out_ch": 3, "num": 1, "type": "mlp", "act": "none"},
            "moe_external_gate": {"in_ch": width, "h_ch": width, "out_ch": width, "num": 2,
                                  "type": "mlp", "act": "none", "out_skip": False},
            "gate_input_norm": {"in_ch": width, "h_ch": 0, "out_ch": 0, "num": 1, "type": "layernorm"},
        },
    }


def make_hparams(num_experts=8, capacity_factor=1.0, bpr=True, model_chunk_size=131072,
                 coarse_samples=257, fine_samples=257, width=256, amp_bf16=False,
                 moe_return_gates=True, appearance_dim=48, nerfmoe_class_name="NeRFMoE",
                 model_cfg=None, **extra):
    """Namespace carrying every hparams field the hot path reads."""
    hp = Namespace(
        model=model_cfg or building_model_cfg(width=width, appearance_dim=appearance_dim),
        nerfmoe_class_name=nerfmoe_class_name,
        moe_capacity_factor=capacity_factor, batch_prioritized_routing=bpr, gate_noise=-1.0,
        compute_balance_loss=False, dispatcher_no_score=False, dispatcher_no_postscore=False,
        moe_expert_type="expertmlp", moe_local_expert_num=num_experts,
        parallel_env=Namespace(global_rank=0), no_expert_parallel=True, single_data_group=None,
        moe_return_gates=moe_return_gates, moe_return_gate_logits=False,
        use_moe_external_gate=True, use_gate_input_norm=True, amp_use_bfloat16=amp_bf16,
        pos_xyz_dim=12, pos_dir_dim=4, appearance_dim=appearance_dim, affine_appearance=False,
        sh_deg=None, shifted_softplus=True,
        # rendering.py fields
        model_chunk_size=model_chunk_size, coarse_samples=coarse_samples, fine_samples=fine_samples,
        perturb=1.0, use_cascade=False, use_sigma_noise=False, sigma_noise_std=1.0,
        white_bkgd=False, use_random_background_color=False, return_pts=False, return_pts_rgb=False,
        return_pts_alpha=False, return_sigma=False, return_alpha=False, use_moe=True,
        bg_use_moe=False, use_load_importance_loss=False, container_path=None, train_mega_nerf=None,
        # mip fields
        weights_resample_padding=0.01, stop_level_grad=True, rgb_padding=0.001,
    )
    for k, v in extra.items():
        setattr(hp, k, v)
    return hp
