"""PPBO settings object -- host-side mirror of the reference's src/ppbo_settings.py:8-79 (same constructor signature and
attribute names, so ppbo_numerical_main.py and the camphor-copper notebook construct it unchanged).  Pure configuration:
nothing here touches the device.  Additive knobs for the B200 path (all default to the reference behaviour) are the
keyword-only arguments after `alpha_grid_distribution`."""

# strategy -> (x acquisition rule, needs the cyclic coordinate counter, needs the cyclic xi-dims list)
_STRATEGIES = {
    "PCD": ("exploit", True, False),
    "EXT": ("exploit", True, False),
    "RAND": ("random", False, False),
    "EI": ("none", False, True),
    "EI-FIXEDX": ("none", False, True),
    "EXR": ("none", False, True),
    "EI-EXT": ("exploit", False, False),
    "EI-EXT-FAST": ("exploit", False, False),
    "EI-VARMAX": ("varmax", False, False),
    "EI-VARMAX-FAST": ("varmax", False, False),
    "COORDINATE-VARMAX": ("varmax", True, False),
}


class PPBO_settings:
    """Settings of Projective Preferential Bayesian Optimization (reference: src/ppbo_settings.py)."""

    def __init__(self, D, bounds, xi_acquisition_function, theta_initial=[1, 0.1, 8], user_feedback_grid_size=100, m=25,
                 verbose=True, EI_EXR_mc_samples=150, EI_EXR_BO_maxiter=20, mustar_finding_trials=3, kernel='SE_kernel',
                 skip_computations_during_initialization=True, skip_xstaroptimization_during_initialization=False,
                 alpha_grid_distribution='equispaced', *, mvn_factor='svd-host', mustar_method='de', mustar_window=16):
        # basic settings (:26-29)
        self.verbose = verbose
        self.user_feedback_grid_size = user_feedback_grid_size
        self.skip_computations_during_initialization = skip_computations_during_initialization
        self.skip_xstaroptimization_during_initialization = skip_xstaroptimization_during_initialization
        # domain (:34-35)
        self.D = D
        self.original_bounds = bounds
        # optimisers (:40-41).  fMAP_optimizer is kept for API parity; the device path always runs its own damped Newton.
        self.fMAP_optimizer = 'trust-exact'
        self.mustar_finding_trials = mustar_finding_trials
        # kernel and hyper-parameters (:44-45)
        self.kernel = kernel
        self.theta_initial = theta_initial
        # pseudo-observations (:48-52)
        self.n_pseudoobservations = m
        self.alpha_grid_distribution = alpha_grid_distribution
        self.TGN_speed = 0.4
        self.n_gausshermite_sample_points = 200
        # acquisition strategy (:56-79)
        self.mc_samples = EI_EXR_mc_samples
        self.BO_maxiter = EI_EXR_BO_maxiter
        self.xi_acquisition_function = xi_acquisition_function
        rule = _STRATEGIES.get(xi_acquisition_function)
        if rule is None:
            print("Unknown acquisition function!")
        else:
            self.x_acquisition_function, cyclic_dim, cyclic_dims = rule
            if cyclic_dim:
                self.dim_query_prev_iter = self.D            # PCD / EXT start from the first coordinate
            if cyclic_dims:
                self.xi_dims_prev_iter = [0, 1] if self.D > 2 else [1]
        # --- B200-path knobs (additive; defaults reproduce the reference) ---
        # 'svd-host': numpy's legacy multivariate_normal factor (bit-compatible draws with the reference for the same RNG
        #             stream); 'device': symmetric eigen-factor computed on the GPU (same distribution, own sign convention)
        self.mvn_factor = mvn_factor
        # 'de': the reference's sequential differential evolution, replayed draw for draw with its loop in C++ (same result bits
        #       as the scipy call, a third of the time); 'de-scipy': the scipy call itself; 'batched': one device evaluation per
        #       generation (scipy 'deferred' updating: another trajectory)
        self.mustar_method = mustar_method
        # trials evaluated per launch by the 'de' search (speculative windows, csrc/de.cu); the result does not depend on it.
        # 16: least host + launch time per retained evaluation (8-12 of 16 trials are retained; at 32 it is 10-14 of 32)
        self.mustar_window = mustar_window
