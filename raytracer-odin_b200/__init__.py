"""raytracer-odin_b200 — host-side mirror of raytracer-odin around the B200 path-tracing library.

Only what the hot path needs:

* ``csrc/``      — the CUDA kernels and the C ABI (``libodinrt_b200.so``, ``include/odinrt_b200.h``)
* ``cabi``       — ctypes mirrors of the C structs
* ``scene``      — the Odin ``Scene`` as numpy arrays + ``finish_scene`` (raytracer.odin:62)
* ``gltf``       — stand-in for ``read_gltf`` (input.odin:13) so scenes can be loaded without Odin
* ``scenegen``   — synthetic glTF generators for the BASELINE.json configs
* ``api``        — ``Renderer``: the call sequence the Odin shim makes through ``foreign import``
* ``output``     — ``get_rgb_image`` / ``save_result`` mirror (output.odin:30,82)
* ``cli``        — ``<gltf> <out> --width --height --ray-depth --num-samples --env-map`` (main.odin:174)

There is no CPU fallback: ``api`` raises if the CUDA library is missing or no GPU is present.
"""

__all__ = ["cabi", "scene", "gltf", "scenegen", "api", "output", "cli"]
__version__ = "0.1.0"
