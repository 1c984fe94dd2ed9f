"""Drop-in mirror of the reference's `aux_code` package for the feature-extraction path: put
`ted-spad_b200/` on sys.path in place of the reference root and
`from aux_code.model_loaders import load_fa_model, load_ft_model` keeps working."""
