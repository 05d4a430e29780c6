"""causalstump_b200 -- B200-native (sm_100a) GP / Student-t-process fitting hot path of
mazphilip/CausalStump behind the reference's own API.

Importing the package does not load CUDA; the first numeric call loads
csrc/libcausalstump_b200.so and raises if it (or a GPU) is missing -- there is no CPU
fallback.
"""
from .main import (CausalStump, CausalStumpObject, KernelClass_GP_SE, KernelClass_TP_SE, check_inputs, list2zero,
                   norm_variables, optAdam, optNadam, optNestorov, predict, treatment)
from .rcpp_exports import (Adam_cpp, Nadam_cpp, Nesterov_cpp, grad_GP_SE_cpp, grad_TP_SE_cpp, invkernel_cpp,
                           kernmat_GP_SE_cpp, kernmat_TP_SE_cpp, mu_solution_cpp, stats_GP_SE, stats_TP_SE)

__all__ = [
    "CausalStump", "predict", "treatment", "CausalStumpObject", "KernelClass_GP_SE", "KernelClass_TP_SE",
    "optAdam", "optNadam", "optNestorov", "list2zero", "norm_variables", "check_inputs",
    "invkernel_cpp", "mu_solution_cpp", "kernmat_GP_SE_cpp", "grad_GP_SE_cpp", "stats_GP_SE",
    "kernmat_TP_SE_cpp", "grad_TP_SE_cpp", "stats_TP_SE", "Nesterov_cpp", "Nadam_cpp", "Adam_cpp",
]
