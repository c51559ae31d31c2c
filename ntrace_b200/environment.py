"""Environment: option registry + `config.conf` reader + command-line overrides, with the reference's key names.

Reference: ``Environment::RegisterOption / ReadEnvFile / Parse`` (src/rt/Environment.cpp:880-1185: file grammar
``Group { key value }`` with ``#`` comments and nested groups; ``-D<Group.key>=<value>`` or ``-<abbrev><value>`` on the
command line; the first positional argument names the environment file) and the option table of
``AppEnvironment::RegisterOptions`` (src/rt/AppEnvironment.cpp:39-174).  Only the knobs that reach the tracing path
are registered with semantics; unknown keys of a reference ``config.conf`` (kd-tree / persistent-builder sections) are
accepted and kept as strings so that shipped configs parse unchanged.
"""
from __future__ import annotations

import re

_BOOL = {"true": True, "on": True, "yes": True, "1": True, "false": False, "off": False, "no": False, "0": False}

# name -> (type, command-line abbreviation, default)   (AppEnvironment.cpp:39-90 + Renderer.numGpus, new)
OPTIONS = {
    "App.benchmark": ("bool", "app_benchmark=", "true"),
    "App.log": ("string", "app_log=", "ntrace.log"),
    "App.stats": ("string", "app_stats=", "stats.log"),
    "App.frameWidth": ("int", "app_frame_width=", "1024"),
    "App.frameHeight": ("int", "app_frame_height=", "768"),
    "Benchmark.scene": ("string", "benchmark_scene=", None),
    "Benchmark.camera": ("string", "benchmark_camera=", None),
    "Benchmark.kernel": ("string", "benchmark_kernel=", None),
    "Benchmark.warmupRepeats": ("int", "benchmark_warmup=", "1"),
    "Benchmark.measureRepeats": ("int", "benchmark_measure=", "5"),
    "Benchmark.cachePath": ("string", "benchmark_cachepath=", None),        # new: directory of the bvhcache files (reference: "bvhcache")
    "Benchmark.pipelined": ("bool", "benchmark_pipelined=", "false"),      # new: queue a frame's batches back to back (Renderer.setPipelined)
    "Benchmark.frameLaunch": ("bool", "benchmark_framelaunch=", "false"),  # new: all batches of a frame in one persistent launch (Renderer.prepareFrame / traceFrame)
    "Renderer.dataStructure": ("string", "renderer_ds=", None),
    "Renderer.builder": ("string", "renderer_builder=", None),
    "Renderer.rayType": ("string", "renderer_raytype=", None),
    "Renderer.samples": ("int", "renderer_samples=", "8"),
    "Renderer.sortRays": ("bool", "renderer_sortrays=", "true"),
    "Renderer.cacheDataStructure": ("bool", "renderer_cache_ds=", "true"),
    "Renderer.numGpus": ("int", "renderer_numgpus=", "1"),
    # new: HLBVHParams of Renderer::getCudaBVH (reference hard-codes {true, 4, 8, 0.001}, Renderer.cpp:201-209) and the SAH-guided collapse
    "HLBVH.bits": ("int", "hlbvh_bits=", "4"),
    "HLBVH.leafSize": ("int", "hlbvh_leafsize=", "8"),
    "HLBVH.collapse": ("bool", "hlbvh_collapse=", "false"),
    "Raygen.random": ("bool", "raygen_random=", "false"),
    "Raygen.aoRadius": ("float", "raygen_aoradius=", "5.0"),
    "Raygen.coherentOrder": ("bool", "raygen_coherent=", "false"),          # new: nt_raygen_set_order(1), direction-coherent slot order inside tiles of <= 1024 rays
    "SBVH.alpha": ("float", "sbvh_alpha=", "1.0e-5"),
}


class EnvironmentError_(RuntimeError):
    pass


class Environment:
    _singleton = None

    def __init__(self):
        self.values = {k: v[2] for k, v in OPTIONS.items() if v[2] is not None}
        self.env_file = None

    # ---- singleton, as the reference uses it
    @classmethod
    def SetSingleton(cls, env):
        cls._singleton = env

    @classmethod
    def GetSingleton(cls):
        if cls._singleton is None:
            cls._singleton = Environment()
        return cls._singleton

    # ---- typed getters (GetIntValue / GetFloatValue / GetBoolValue / GetStringValue)
    def _raw(self, name):
        if name not in self.values:
            raise EnvironmentError_(f"Environment: option {name} is not set")
        return self.values[name]

    def GetString(self, name) -> str:
        return str(self._raw(name))

    def GetInt(self, name) -> int:
        return int(self._raw(name))

    def GetFloat(self, name) -> float:
        return float(self._raw(name))

    def GetBool(self, name) -> bool:
        v = str(self._raw(name)).strip().lower()
        if v not in _BOOL:
            raise EnvironmentError_(f"Environment: option {name} has a non-boolean value '{v}'")
        return _BOOL[v]

    def Has(self, name) -> bool:
        return name in self.values

    def Set(self, name, value):
        t = OPTIONS.get(name, ("string",))[0]
        v = str(value)
        try:
            if t == "int":
                int(v)
            elif t == "float":
                float(v)
            elif t == "bool" and v.strip().lower() not in _BOOL:
                raise ValueError
        except ValueError:
            raise EnvironmentError_(f"Environment: bad value '{v}' for {t} option {name}") from None
        self.values[name] = v

    # ---- environment file: Group { key value ... } with # comments, nested groups allowed
    def ReadEnvFile(self, path: str):
        with open(path, "r", errors="replace") as f:
            self.ParseEnvString(f.read(), path)
        self.env_file = path

    def ParseEnvString(self, text: str, origin: str = "<string>"):
        prefix = []
        for lineno, raw in enumerate(text.splitlines(), 1):
            line = raw.split("#", 1)[0]
            toks = re.findall(r"\{|\}|[^\s{}]+", line)
            i = 0
            while i < len(toks):
                tok = toks[i]
                if tok == "}":
                    if not prefix:
                        raise EnvironmentError_(f"Error: unpaired }} in {origin} (line {lineno}).")
                    prefix.pop()
                    i += 1
                elif i + 1 < len(toks) and toks[i + 1] == "{":
                    prefix.append(tok)
                    i += 2
                elif tok == "{":
                    raise EnvironmentError_(f"Error: group without a name in {origin} (line {lineno}).")
                else:
                    # key value: the value is the rest of the line up to a brace
                    j = i + 1
                    val = []
                    while j < len(toks) and toks[j] not in "{}":
                        val.append(toks[j]); j += 1
                    if not val:
                        raise EnvironmentError_(f"Error: option {'.'.join(prefix + [tok])} has no value in {origin} (line {lineno}).")
                    self.Set(".".join(prefix + [tok]), " ".join(val))
                    i = j
        if prefix:
            raise EnvironmentError_(f"Error: unclosed group {'.'.join(prefix)} in {origin}.")

    # ---- command line: [envfile] -D<Group.key>=<v> | -<abbrev><v>
    def Parse(self, argv, default_env_file: str | None = None):
        args = list(argv)
        env_file = default_env_file
        rest = []
        for a in args:
            if not a.startswith("-") and env_file is default_env_file and not rest:
                env_file = a
            else:
                rest.append(a)
        if env_file:
            self.ReadEnvFile(env_file)
        for a in rest:
            if a.startswith("-D"):
                if "=" not in a:
                    raise EnvironmentError_(f"Environment: malformed option '{a}'")
                k, v = a[2:].split("=", 1)
                self.Set(k, v)
                continue
            body = a.lstrip("-")
            for name, (_, abbrev, _) in OPTIONS.items():
                if abbrev and body.startswith(abbrev):
                    self.Set(name, body[len(abbrev):])
                    break
            else:
                raise EnvironmentError_(f"Environment: unknown option '{a}'")
        return True
