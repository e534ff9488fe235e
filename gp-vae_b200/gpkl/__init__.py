"""gpkl -- B200-native GP-prior KL path for GP-VAE (host side over the C ABI of include/gpkl.h)."""
from .ops import (GpPriorKL, gp_prior_kl, gp_prior_kl_forward, gp_prior_kl_backward, HostStep,
                  workspace_bytes, bernoulli_recon, elbo_loss, gp_recog_sample, collate_batch, gp_posterior_impute,
                  release_workspaces, philox_normal)
from .reference_api import (GPPriorPath, GPRecogPath, SyntheticDataHandlerGPU, sample_given_part_latent,
                            post_gp_sample)
from . import _lib

__all__ = ["GpPriorKL", "gp_prior_kl", "gp_prior_kl_forward", "gp_prior_kl_backward", "HostStep",
           "workspace_bytes", "GPPriorPath", "GPRecogPath", "SyntheticDataHandlerGPU", "_lib", "bernoulli_recon", "elbo_loss", "gp_recog_sample", "collate_batch",
           "gp_posterior_impute", "release_workspaces", "philox_normal", "sample_given_part_latent", "post_gp_sample"]
