"""Drop-in mirror of the reference's `layers/` package (module names, class names, constructor argument order,
forward signatures, attributes and state_dict keys are the reference's; the bodies call the sm_100a kernels).
See INTEGRATION.md for how `models/GNNs.py` picks these up unchanged."""
