"""``python setup.py build_ext --inplace`` — same command as the reference (README.md:107,
/root/reference/setup.py:9-47), but the extension is built by nvcc for sm_100a instead of CMake."""
from setuptools import Extension, setup
from setuptools.command.build_ext import build_ext


class NvccBuild(build_ext):
    def run(self):
        import importlib.util
        from pathlib import Path

        spec = importlib.util.spec_from_file_location("fj_build", Path(__file__).parent / "flash_hash_join_b200" / "build.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build(force=bool(self.force), verbose=True)


setup(
    name="flash_hash_join_b200",
    version="0.1.0",
    description="B200-native (sm_100a) equi-join engine behind flash_join's Python API",
    packages=["flash_hash_join_b200"],
    ext_modules=[Extension("flash_hash_join_b200.flash_join", sources=[])],
    cmdclass={"build_ext": NvccBuild},
    zip_safe=False,
)
