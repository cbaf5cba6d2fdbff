"""`pcdet`-shaped boundary of the B200 CAGroup3D path.

Only the names tools/test.py touches for CAGroup3D exist: config, datasets.build_dataloader,
models.build_network / load_data_to_gpu, the four registries, utils.common_utils.  The classes behind
the registries live in cagroup3d_b200/ and call the CUDA C ABI; there is no CPU fallback.
"""
__version__ = "0.5.2+b200"
