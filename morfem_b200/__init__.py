"""morfem_b200 -- B200-native (sm_100a) implementation of morfem's reduced-order frequency-sweep hot path.

Public surface mirrors the reference's two modules:

    from morfem_b200.implementation import morfem, ModelDefinition, solve_finite_element_method
    from morfem_b200.test_helpers import finite_element_method_model_order_reduction_gsm, b_coefficient

Importing this package does not load CUDA; the first call into a hot stage loads ``libmorfem_b200.so`` and
raises if it (or a GPU) is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"
__all__ = ["implementation", "test_helpers", "device", "synthetic", "dist"]
