"""updes_b200 -- B200-native global RBF-collocation hot path of ddrous/Updes.

Drop-in for the path  pde_solver / pde_solver_jit  over a SquareCloud or GmshCloud with the built-in
kernels, max_degree polynomial augmentation and Dirichlet / Neumann / Robin / periodic facets
(reference: updes/operators.py:559-683, updes/assembly.py).  Compute runs in hand-written CUDA for
sm_100a behind the C-ABI of include/updes_b200.h; there is no CPU fallback.
"""
from .cloud import Cloud, GmshCloud, SquareCloud
from .rbf import (compute_nb_monomials, distance, gaussian, gaussian_func, identify_rbf, inv_multiquadric_func, inverse_multiquadric,
                  make_all_monomials, make_monomial, make_nodal_rbf, multiquadric, multiquadric_func, polyharmonic, polyharmonic_func,
                  thin_plate, thin_plate_func)
from .operators import (apply_neumann_conditions, cartesian_gradient, cartesian_gradient_vec, enforce_cartesian_gradient_neumann,
                        BatchPoints, OperatorLoweringError, SteadySol, assemble_q, boundary_conditions_func_to_arr, clear_cache,
                        disable_distributed, enable_distributed, integrate_field, interpolate_field,
                        compute_coefficients, core_compute_coefficients, divergence, divergence_vec, dot,
                        duplicate_robin_coeffs, get_field_coefficients, gradient, gradient_vals, gradient_vals_vec, gradient_vec,
                        laplacian, laplacian_vals, laplacian_vals_vec,
                        laplacian_vec, lower_diff_operator, nodal_div_grad, nodal_gradient, nodal_laplacian,
                        nodal_value, pde_multi_solver, pde_solver, pde_solver_jit, pde_solver_jit_with_bc, value, value_vec,
                        zerofy_periodic_cond)

from .operators import value_vec_, gradient_vec_  # noqa: E402  (operators.py:149, :183: the un-jitted names)
from .operators import gradient_vals_vec as gradient_vals_vec_, laplacian_vals_vec as laplacian_vals_vec_  # noqa: E402
from .autodiff import linear_solve
from .utils import RK4, dataloader, dot_mat, dot_vec, make_dir, plot, print_line_by_line, random_name
# the reference's demo scripts take these names from `from updes import *` (updes/utils.py imports them at module level);
# Partial is jax.tree_util.Partial there -- a partial application, which is all the demos use it for
import os  # noqa: E402,F401
from functools import partial  # noqa: E402,F401
from functools import partial as Partial  # noqa: E402,F401
try:                                        # the reference's scripts also pick jax / jnp up from the star import
    import jax  # noqa: E402,F401
    import jax.numpy as jnp  # noqa: E402,F401
except ImportError:                         # JAX is optional here (absent from this image): the product computes with numpy + CUDA
    pass
try:                                        # plt / sns reach the demos through the star import too; both are optional here
    import matplotlib.pyplot as plt  # noqa: E402,F401
except ImportError:
    pass
try:
    import seaborn as sns  # noqa: E402,F401
except ImportError:
    pass
from .explicit import (assemble_A, assemble_B, assemble_P, assemble_Phi, assemble_bd_Phi_P, assemble_invert_A,
                       assemble_op_Phi_P)

__version__ = "0.1.0"
