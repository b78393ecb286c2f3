"""ctypes binding of the C ABI declared in include/tonic_b200.h.

The product path has NO CPU fallback: if the CUDA library is missing or a call
fails, an exception is raised.
"""

import ctypes
import os

c_int, c_i32, c_i64 = ctypes.c_int, ctypes.c_int32, ctypes.c_int64
c_f, c_d, c_vp = ctypes.c_float, ctypes.c_double, ctypes.c_void_p
c_u32, c_u64 = ctypes.c_uint32, ctypes.c_uint64

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libtonic_b200.so')

ACT_TANH, ACT_RELU = 0, 1

STAT_ROWS, STAT_LOSS, STAT_KL, STAT_ENTROPY, STAT_CLIPPED = 0, 1, 2, 3, 4
STAT_NONZERO_ADV, STAT_STD, STAT_VALUE, STAT_VALUE2 = 5, 6, 7, 8
STAT_COUNT = 12


class TbEnv(ctypes.Structure):
    _fields_ = [('n_envs', c_i32), ('obs_dim', c_i32), ('act_dim', c_i32),
                ('max_episode_steps', c_i32), ('seed', c_i64), ('first_worker', c_i64),
                ('d_state', c_vp), ('d_length', c_vp), ('d_episode', c_vp),
                ('d_score', c_vp), ('d_ep_scores', c_vp), ('d_ep_lengths', c_vp),
                ('d_ep_count', c_vp), ('log_cap', c_i32), ('time_feature', c_i32),
                ('time_low', c_f), ('time_high', c_f), ('task', c_i32), ('d_state64', c_vp)]


class TbMlpShape(ctypes.Structure):
    _fields_ = [('d_in', c_i32), ('hidden', c_i32), ('n_out', c_i32), ('act', c_i32),
                ('off_w1', c_i32), ('off_b1', c_i32), ('off_w2', c_i32), ('off_b2', c_i32),
                ('off_w3', c_i32), ('off_b3', c_i32), ('n_params', c_i32),
                ('off_w1t', c_i32), ('off_w2t', c_i32), ('n_packed', c_i32),
                ('off_w2_hi', c_i32), ('off_w2_lo', c_i32), ('off_w2t_hi', c_i32),
                ('off_w2t_lo', c_i32), ('off_w1_img_hi', c_i32), ('off_w1_img_lo', c_i32)]


class TbMlpInput(ctypes.Structure):
    _fields_ = [('d_x1', c_vp), ('dim1', c_i32), ('d_mean', c_vp), ('d_std', c_vp),
                ('d_x2', c_vp), ('dim2', c_i32), ('gather2', c_i32), ('d_idx', c_vp)]


class TbPeers(ctypes.Structure):
    _fields_ = [('world', c_i32), ('rank', c_i32), ('base', c_vp * 8)]


class TbAdam(ctypes.Structure):
    _fields_ = [('lr', c_d), ('beta1', c_d), ('beta2', c_d), ('eps', c_d),
                ('n_params', c_i32), ('d_params', c_vp), ('d_m', c_vp), ('d_v', c_vp),
                ('d_step', c_vp)]


_P = ctypes.POINTER

_PROTOTYPES = {
    'tb_version': (c_int, []),
    'tb_last_error': (ctypes.c_char_p, []),
    'tb_launch_count': (c_i64, []),
    'tb_env_start': (c_int, [_P(TbEnv), c_vp, c_vp]),
    'tb_env_step': (c_int, [_P(TbEnv), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tb_act_env_step': (c_int, [_P(TbEnv), c_vp, c_vp, c_u64, c_u64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                c_vp, c_vp, c_vp]),
    'tb_rollout_fused': (c_int, [_P(TbEnv), _P(TbMlpShape), c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp,
                                 c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_u64, c_u64, c_u64, c_vp, c_vp]),
    'tb_moments_record': (c_int, [c_vp, c_i64, c_i32, c_vp, c_vp]),
    'tb_moments_update': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_f, c_vp]),
    'tb_lambda_returns': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_d, c_d, c_vp]),
    'tb_advantages': (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i32, c_vp]),
    'tb_mlp_forward': (c_int, [_P(TbMlpShape), c_vp, c_vp, _P(TbMlpInput), c_i64, c_vp, c_vp,
                               c_vp, c_vp, c_vp, c_vp]),
    'tb_mlp_backward': (c_int, [_P(TbMlpShape), c_vp, c_vp, c_i32, c_vp, c_vp, c_i64, c_vp, c_vp,
                                c_vp, c_i32, c_i32, c_vp, c_vp]),
    'tb_mlp_wgrad': (c_int, [_P(TbMlpShape), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32,
                             c_i32, c_i64, c_vp, c_i32, c_vp, c_vp]),
    'tb_adam_step': (c_int, [_P(TbAdam), _P(TbMlpShape), c_vp, c_vp, c_i32, c_i32, c_f, c_vp, c_vp,
                             c_f, c_vp, c_vp]),
    'tb_reduce_partials': (c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    'tb_grad_sqnorm': (c_int, [c_vp, c_i32, c_vp, c_vp, c_vp]),
    'tb_grad_clip': (c_int, [c_vp, c_i32, c_vp, c_f, c_f, c_vp, c_vp]),
    'tb_peer_region_bytes': (c_i64, [c_i32]),
    'tb_peer_region_bytes_fused': (c_i64, [c_i32]),
    'tb_peer_publish': (c_int, [_P(TbPeers), c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp,
                                c_vp, c_vp]),
    'tb_adam_step_peers': (c_int, [_P(TbAdam), _P(TbMlpShape), c_vp, _P(TbPeers), c_f, c_vp, c_vp, c_vp,
                                   c_i32, c_f, c_vp, c_vp]),
    'tb_mlp_pack': (c_int, [_P(TbMlpShape), c_vp, c_vp, c_vp]),
    'tb_soft_update': (c_int, [c_vp, c_vp, c_i64, c_d, c_vp]),
    'tb_gauss_sample': (c_int, [c_vp, c_vp, c_vp, c_u64, c_u64, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    'tb_gauss_policy_loss': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_f, c_f,
                                     c_vp, c_vp, c_vp, c_vp]),
    'tb_mse_loss': (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp]),
    'tb_tanh_action': (c_int, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_u64, c_u64, c_f, c_f, c_vp, c_vp]),
    'tb_squashed_sample': (c_int, [c_vp, c_vp, c_u64, c_u64, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    'tb_q_target': (c_int, [c_vp, c_vp, c_vp, c_d, c_vp, c_vp, c_vp, c_d, c_i64, c_vp, c_vp]),
    'tb_q_actor_loss': (c_int, [c_vp, c_vp, c_vp, c_d, c_i64, c_vp, c_vp, c_vp, c_vp]),
    'tb_dpg_head_grad': (c_int, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    'tb_sac_head_grad': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_d, c_i64, c_i32, c_vp, c_vp]),
    'tb_split_tf32': (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    'tb_tc_gemm256': (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp,
                              c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp]),
    'tb_mlp_forward_tc': (c_int, [_P(TbMlpShape), c_vp, c_vp, _P(TbMlpInput), c_i64, c_vp, c_vp, c_vp,
                                  c_vp, c_vp, c_i32, c_vp, c_vp]),
    'tb_tc_mlp_train': (c_int, [_P(TbMlpShape), c_vp, c_vp, _P(TbMlpInput), c_i64, c_i32, c_vp, c_vp, c_vp,
                                c_vp, c_vp, c_vp, c_f, c_f, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32,
                                c_vp, c_vp, c_vp, c_i32, c_vp, c_vp]),
    'tb_debug_plain_hi': (c_int, [c_i32]),
    'tb_debug_skinny': (c_int, [c_i32]),
    'tb_tc_timeline': (c_int, [c_vp]),
    'tb_wgrad_timeline': (c_int, [c_vp]),
    'tb_q_target_discounts': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_d, c_i64, c_vp, c_vp]),
    'tb_replay_accumulate_n_steps': (c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32,
                                             c_i32, c_vp]),
    'tb_tc_mlp_forward_vloss': (c_int, [_P(TbMlpShape), c_vp, c_vp, _P(TbMlpInput), c_i64, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_i32,
                                        c_i32, c_vp, c_vp]),
    'tb_tc_mlp_backward': (c_int, [_P(TbMlpShape), c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_i64,
                                   c_vp, c_vp, c_vp, c_i32, c_vp, c_vp]),
    'tb_tc_mlp_forward': (c_int, [_P(TbMlpShape), c_vp, c_vp, _P(TbMlpInput), c_i64, c_vp, c_vp, c_vp,
                                  c_vp, c_vp, c_i32, c_vp, c_vp]),
    'tb_mlp_backward_tc': (c_int, [_P(TbMlpShape), c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_i64,
                                   c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    'tb_mlp_wgrad_tc': (c_int, [_P(TbMlpShape), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32,
                                c_i32, c_i32, c_i64, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    'tb_mlp_wgrad_fused': (c_int, [_P(TbMlpShape), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32,
                                   c_i32, c_i32, c_i64, c_vp, c_i32, c_vp, c_vp, c_i32, _P(TbAdam), c_vp,
                                   c_f, c_vp, c_f, c_vp, c_vp, _P(TbPeers), c_vp, c_vp, c_vp]),
    'tb_tc_wgrad256': (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp,
                               c_vp]),
    'tb_permutation': (c_int, [c_u64, c_u64, c_vp, c_i64, c_vp, c_vp]),
    'tb_counter_add': (c_int, [c_vp, c_u64, c_vp]),
    'tb_array_stats': (c_int, [c_vp, c_i64, c_vp, c_vp]),
    'tb_set_noise_base': (c_int, [c_vp]),
    'tb_randint': (c_int, [c_u64, c_u64, c_vp, c_vp, c_i64, c_vp, c_vp]),
    'tb_ring_store': (c_int, [_P(c_vp), _P(c_vp), _P(c_i64), c_i32, c_vp, c_vp]),
    'tb_ring_advance': (c_int, [c_vp, c_i64, c_i64, c_vp]),
    'tb_profile_begin': (c_int, []),
    'tb_profile_end': (c_int, [ctypes.c_char_p, c_i32]),
    'tb_rs_create': (c_vp, [c_u32]),
    'tb_rs_destroy': (None, [c_vp]),
    'tb_rs_shuffle_i64': (None, [c_vp, c_vp, c_i64]),
    'tb_rs_randint': (None, [c_vp, c_i64, c_vp, c_i64]),
    'tb_rs_uniform': (None, [c_vp, c_d, c_d, c_vp, c_i64]),
    'tb_rs_normal': (None, [c_vp, c_vp, c_i64]),
}

# entry points whose int return value is a status code
_CHECKED = {name for name, (res, _) in _PROTOTYPES.items()
            if res is c_int and name not in ('tb_version', 'tb_profile_end')}
_PROTOTYPES_DONE = True

_lib = None


class TonicB200Error(RuntimeError):
    pass


def load():
    """Loads libtonic_b200.so (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TonicB200Error(
            f'{LIB_PATH} not found: build it with `python -m tonic_b200.build` '
            '(or __graft_entry__.build()); there is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_PROTOTYPES)


def call(name, *args):
    """Calls an entry point and raises TonicB200Error on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _CHECKED and rc != 0:
        msg = lib.tb_last_error().decode(errors='replace')
        raise TonicB200Error(f'{name} failed with status {rc}: {msg}')
    return rc


def launch_count():
    return int(load().tb_launch_count())


def ptr(tensor):
    """Device (or host) address of a torch tensor / numpy array, or None."""
    if tensor is None:
        return None
    if hasattr(tensor, 'data_ptr'):
        return tensor.data_ptr()
    return tensor.ctypes.data


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
