"""Test-only oracle for the Diff-VITS denoising loop (see unet_oracle.py / sampler_oracle.py headers).
Nothing under diff-vits_b200/ imports this package."""
