"""`pip install -e .` (or `python setup.py build_ext`) compiles the sm_100a extension IN-TREE
(harmony_b200/ops/_build/_hb_native*.so) through harmony_b200.ops.build -- the same path __graft_entry__.build()
uses -- so the shared object sits next to the package and is loaded by path, never from a JIT cache."""
from setuptools import Command, setup
from setuptools.command.build_py import build_py


class build_ext(Command):
    description = "compile the sm_100a CUDA/C++ extension in-tree"
    user_options = [("force", "f", "rebuild even if up to date")]

    def initialize_options(self):
        self.force = False

    def finalize_options(self):
        pass

    def run(self):
        from harmony_b200.ops import build
        build.build(verbose=True, force=bool(self.force))


class build_py_with_ext(build_py):
    def run(self):
        try:
            self.run_command("build_ext")
        except Exception as e:  # noqa: BLE001  (CPU-only checkouts still install the Python runtime)
            print(f"warning: native extension not built ({e}); host backend only")
        super().run()


setup(cmdclass={"build_ext": build_ext, "build_py": build_py_with_ext})
