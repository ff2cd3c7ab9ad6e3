import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import importlib.util
spec = importlib.util.spec_from_file_location("qt", "tools/quick_time.py"); qt = importlib.util.module_from_spec(spec); spec.loader.exec_module(qt)
qt.run(1, "pinn", "FBC", 1 << 20, iters=10)
qt.run(2, "drm", "RB", 1 << 20, iters=10)
qt.run(2, "pinn", "FBC", 1 << 20, iters=10)
qt.run(3, "pinn", "FBC", 1 << 22, iters=5)
qt.run(5, "drm", "RB", 1 << 20, iters=10)
