"""ctypes binding of libtqf.so (C ABI declared in include/tqf.h).

The library is the product: there is no Python / CPU fallback.  Importing this
module fails loudly when the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or
`make -C tf-quant-finance_b200/csrc`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TQF_LIBRARY: developer override used to A/B differently tuned builds.
LIB_PATH = os.environ.get('TQF_LIBRARY') or os.path.join(_HERE, 'lib', 'libtqf.so')

TQF_OK = 0
TQF_ERR_INVALID_ARGUMENT = -1
TQF_ERR_UNSUPPORTED = -2
TQF_ERR_CUDA = -3
TQF_ERR_IO = -4

F32, F64 = 0, 1
RNG_PHILOX, RNG_SOBOL, RNG_DRAWS = 1, 2, 3
MODEL_AFFINE_1F = 1
MODEL_GBM_1F = 2
MODEL_HESTON_EULER = 3
MODEL_HESTON_QE = 4
MODEL_MVGBM = 5
MODEL_LINEAR_1F = 6
MODEL_HW1F = 7
MODEL_AFFINE_ND = 8
MODEL_AFFINE_1F_TANGENT = 9
MODEL_MILSTEIN_1F = 10
MODEL_HJM = 11
MODEL_HESTON_TANGENT = 12
PAYOFF_CALL = 1
PAYOFF_PUT = 2
PAYOFF_UP_OUT_CALL = 3
PAYOFF_DOWN_OUT_PUT = 4
PAYOFF_UP_OUT_PUT = 5
PAYOFF_DOWN_OUT_CALL = 6
PAYOFF_IDENTITY = 7
PAYOFF_HW_SWAPTION = 8
PAYOFF_CALL_TANGENT = 9
PAYOFF_PUT_TANGENT = 10
TRANSFORM_NONE, TRANSFORM_EXP = 0, 1
MAX_PAYOFFS = 8
MAX_SWAPTION_PAYMENTS = 64


class RngDesc(C.Structure):
  _fields_ = [
      ('type', C.c_int32),
      ('antithetic', C.c_int32),
      ('key', C.c_uint32 * 2),
      ('counter', C.c_uint32 * 4),
      ('skip', C.c_uint64),
      ('direction_numbers', C.c_void_p),
      ('draws_dev', C.c_void_p),
      ('unit_stride', C.c_uint64),
      ('unit_offset', C.c_uint64),
  ]


class ModelDesc(C.Structure):
  _fields_ = [
      ('kind', C.c_int32),
      ('dtype', C.c_int32),
      ('dim', C.c_int32),
      ('num_factors', C.c_int32),
      ('num_steps', C.c_int32),
      ('num_steps_total', C.c_int32),
      ('num_coef', C.c_int32),
      ('reserved', C.c_int32),
      ('coef', C.c_void_p),
      ('x0', C.c_void_p),
      ('matrix', C.c_void_p),
      ('vector', C.c_void_p),
      ('x0_paths_dev', C.c_void_p),
  ]


class PayoffDesc(C.Structure):
  _fields_ = [
      ('kind', C.c_int32),
      ('component', C.c_int32),
      ('transform', C.c_int32),
      ('tangent_component', C.c_int32),
      ('strike', C.c_double),
      ('barrier', C.c_double),
      ('scale', C.c_double),
      ('expiry_step', C.c_int32),
      ('num_payments', C.c_int32),
      ('is_payer', C.c_int32),
      ('brownian_bridge', C.c_int32),
      ('num_factors', C.c_int32),
      ('reserved3', C.c_int32),
      ('reserved4', C.c_double),
      ('pay_g', C.c_double * MAX_SWAPTION_PAYMENTS),
      ('pay_k', C.c_double * MAX_SWAPTION_PAYMENTS),
      ('pay_coef', C.c_double * MAX_SWAPTION_PAYMENTS),
  ]


class LsmDesc(C.Structure):
  _fields_ = [
      ('dtype', C.c_int32),
      ('dim', C.c_int32),
      ('batch', C.c_int32),
      ('basis_size', C.c_int32),
      ('exponents', C.c_void_p),
      ('strikes', C.c_void_p),
      ('num_paths', C.c_uint64),
      ('path_offset', C.c_uint64),
      ('num_calibration_samples', C.c_uint64),
      ('paths_dev', C.c_void_p),
      ('stride_path', C.c_int64),
      ('stride_time', C.c_int64),
      ('stride_dim', C.c_int64),
      ('stride_batch', C.c_int64),
      ('w_dev', C.c_void_p),
      ('partials_dev', C.c_void_p),
      ('partials_doubles', C.c_uint64),
      ('exercise_time_indices', C.c_void_p),
      ('num_exercise_times', C.c_int32),
      ('reserved', C.c_int32),
      ('exercise_values_dev', C.c_void_p),
      ('path_ratio_dev', C.c_void_p),
  ]


class TqfError(RuntimeError):
  """A libtqf call failed (status code + tqf_last_error())."""

  def __init__(self, code, message):
    super().__init__('libtqf error {}: {}'.format(code, message))
    self.code = code


# The symbols include/tqf.h declares: name -> (restype, argtypes).
_u32p = C.POINTER(C.c_uint32)
_SIGNATURES = {
    'tqf_last_error': (C.c_char_p, []),
    'tqf_version': (C.c_int, []),
    'tqf_device_count': (C.c_int, []),
    'tqf_abi_sizes': (C.c_int, [C.POINTER(C.c_int32)]),
    'tqf_philox_stateless_key_counter':
        (C.c_int, [C.POINTER(C.c_int64), _u32p, _u32p]),
    'tqf_philox_stateful_key_counter': (C.c_int, [C.c_int64, _u32p, _u32p]),
    'tqf_philox_raw_fill':
        (C.c_int, [_u32p, _u32p, C.c_uint64, C.c_uint64, C.c_void_p,
                   C.c_void_p]),
    'tqf_philox_normal_fill':
        (C.c_int, [_u32p, _u32p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p,
                   C.c_void_p]),
    'tqf_philox_uniform_fill':
        (C.c_int, [_u32p, _u32p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p,
                   C.c_void_p]),
    'tqf_sobol_direction_numbers':
        (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                   C.c_void_p]),
    'tqf_sobol_direction_numbers_from_file':
        (C.c_int, [C.c_char_p, C.c_int, C.c_void_p]),
    'tqf_sobol_fill':
        (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                   C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_philox_uniform_int_fill':
        (C.c_int, [_u32p, _u32p, C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_qmc_sobol_generating_matrices':
        (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                   C.c_void_p]),
    'tqf_qmc_scramble_generating_matrices':
        (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'tqf_qmc_digital_net_fill':
        (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                   C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_qmc_lattice_rule_fill':
        (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                   C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_plan_create':
        (C.c_int, [C.POINTER(ModelDesc), C.POINTER(RngDesc), C.c_uint64,
                   C.POINTER(C.c_void_p)]),
    'tqf_plan_destroy': (C.c_int, [C.c_void_p]),
    'tqf_plan_price':
        (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(PayoffDesc),
                   C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_plan_price_host':
        (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(PayoffDesc),
                   C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'tqf_plan_paths':
        (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                   C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    'tqf_plan_set_peer_exchange':
        (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_uint64]),
    'tqf_plan_peer_epoch': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    'tqf_plan_set_sobol_clamp': (C.c_int, [C.c_void_p, C.c_int]),
    'tqf_lsm_persistent_eligible': (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    'tqf_lsm_run_persistent':
        (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_double,
                   C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'tqf_lsm_status': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    'tqf_hw_exercise_values':
        (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_hw_discount_curves':
        (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_hjm_discount_curves':
        (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                   C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_plan_paths_sums':
        (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                   C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_halton_fill':
        (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                   C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_halton_permutations':
        (C.c_int, [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'tqf_halton_randomized_fill':
        (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                   C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_lsm_workspace': (C.c_int, [C.POINTER(LsmDesc), C.c_int, C.POINTER(C.c_uint64)]),
    'tqf_lsm_create': (C.c_int, [C.POINTER(LsmDesc), C.POINTER(C.c_void_p)]),
    'tqf_lsm_destroy': (C.c_int, [C.c_void_p]),
    'tqf_lsm_column_sums':
        (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'tqf_lsm_set_fused_solve':
        (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    'tqf_lsm_run_fused':
        (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                   C.c_void_p, C.c_void_p]),
    'tqf_lsm_fused_eligible': (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    'tqf_lsm_peer_bytes': (C.c_int, [C.POINTER(C.c_uint64)]),
    'tqf_lsm_set_peer_exchange':
        (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_uint64]),
    'tqf_lsm_peer_epoch': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    'tqf_peer_alloc': (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p]),
    'tqf_peer_open': (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    'tqf_peer_close': (C.c_int, [C.c_void_p]),
    'tqf_peer_free': (C.c_int, [C.c_void_p]),
    'tqf_peer_status': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    'tqf_lsm_init': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    'tqf_lsm_step':
        (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                   C.c_void_p]),
    'tqf_lsm_solve':
        (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    'tqf_lsm_sums_layout':
        (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'tqf_lsm_value_sum': (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    'tqf_math_eval':
        (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    'tqf_measure_fp64_peak':
        (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))

_lib = None


def lib():
  """Loads libtqf.so once; raises ImportError if it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise ImportError(
          'libtqf.so not found at {}. Build it with `make -C '
          'tf-quant-finance_b200/csrc` (or __graft_entry__.build()); there is '
          'no Python fallback.'.format(LIB_PATH))
    handle = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
      fn = getattr(handle, name)
      fn.restype = restype
      fn.argtypes = argtypes
    _lib = handle
  return _lib


def check(code):
  if code != TQF_OK:
    msg = lib().tqf_last_error()
    msg = msg.decode('utf-8', 'replace') if msg else ''
    if code == TQF_ERR_INVALID_ARGUMENT:
      raise ValueError('libtqf: ' + msg)
    if code == TQF_ERR_UNSUPPORTED:
      raise NotImplementedError('libtqf: ' + msg)
    raise TqfError(code, msg)


def require_cuda():
  """Fails loudly when there is no CUDA device (no CPU fallback)."""
  if lib().tqf_device_count() <= 0:
    raise TqfError(TQF_ERR_CUDA,
                   'no CUDA device visible: the B200 engine has no CPU '
                   'fallback')
