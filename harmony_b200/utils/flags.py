"""Flag / configuration system.

The reference configures everything through Apache Tang named parameters with a CLI short name
(`-short_name value`), immutable mergeable configurations that serialise to strings (that is how
a whole job description travels over the job-server socket) and per-destination extraction
(client / driver / master / server / worker / user).  See SURVEY.md 5.6 and
common/.../param/Parameters.java:25-119, utils/ConfigurationUtils.java.

Here: `Param` declares one named parameter; `ParamRegistry` parses `-name value` command lines
(unknown flags are an error unless `allow_unknown`); `Configuration` is an immutable mapping that
merges, extracts by parameter group and serialises to JSON.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Any, Callable, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple


def _parse_bool(s: Any) -> bool:
    if isinstance(s, bool):
        return s
    v = str(s).strip().lower()
    if v in ("true", "1", "yes", "y", "on"):
        return True
    if v in ("false", "0", "no", "n", "off"):
        return False
    raise ValueError(f"not a boolean: {s!r}")


_PARSERS: Dict[type, Callable[[Any], Any]] = {bool: _parse_bool, int: lambda s: int(float(s)) if isinstance(s, str) and "." in s else int(s),
                                              float: float, str: str}


@dataclass(frozen=True)
class Param:
    """One named parameter. `name` is the CLI short name (flag is `-name`)."""
    name: str
    type: type = str
    default: Any = None
    doc: str = ""
    group: str = "user"
    required: bool = False

    def parse(self, raw: Any) -> Any:
        if raw is None:
            return None
        return _PARSERS.get(self.type, self.type)(raw)


class Configuration(Mapping[str, Any]):
    """Immutable name -> value mapping (the Tang `Configuration` equivalent)."""

    def __init__(self, values: Optional[Mapping[str, Any]] = None) -> None:
        self._values: Dict[str, Any] = dict(values or {})

    def __getitem__(self, key: str) -> Any:
        return self._values[key]

    def __iter__(self):
        return iter(self._values)

    def __len__(self) -> int:
        return len(self._values)

    def get(self, key: str, default: Any = None) -> Any:  # type: ignore[override]
        return self._values.get(key, default)

    def merge(self, *others: Mapping[str, Any]) -> "Configuration":
        out = dict(self._values)
        for o in others:
            out.update(dict(o))
        return Configuration(out)

    def with_values(self, **kv: Any) -> "Configuration":
        return self.merge(kv)

    def extract(self, params: Iterable[Param | str]) -> "Configuration":
        names = [p.name if isinstance(p, Param) else p for p in params]
        return Configuration({n: self._values[n] for n in names if n in self._values})

    def serialize(self) -> str:
        return json.dumps(self._values, sort_keys=True, default=str)

    @staticmethod
    def deserialize(s: str) -> "Configuration":
        return Configuration(json.loads(s))

    def __repr__(self) -> str:
        return f"Configuration({self._values!r})"


class FlagError(ValueError):
    """A bad, unknown or missing CLI flag."""

    pass


class ParamRegistry:
    """Registered short names -> Param; parses `-name value` argv (CommandLine equivalent)."""

    def __init__(self, params: Iterable[Param] = ()) -> None:
        self._params: Dict[str, Param] = {}
        for p in params:
            self.register(p)

    def register(self, p: Param) -> "ParamRegistry":
        old = self._params.get(p.name)
        if old is not None and old != p:
            # same flag registered by two components must agree on type
            if old.type != p.type:
                raise FlagError(f"flag -{p.name} registered twice with different types")
        self._params[p.name] = p
        return self

    def register_all(self, ps: Iterable[Param]) -> "ParamRegistry":
        for p in ps:
            self.register(p)
        return self

    def params(self) -> List[Param]:
        return list(self._params.values())

    def by_group(self, group: str) -> List[Param]:
        return [p for p in self._params.values() if p.group == group]

    def __contains__(self, name: str) -> bool:
        return name in self._params

    def defaults(self) -> Configuration:
        return Configuration({p.name: p.default for p in self._params.values() if p.default is not None})

    def parse(self, argv: Sequence[str], allow_unknown: bool = False,
              apply_defaults: bool = True) -> Tuple[Configuration, List[str]]:
        values: Dict[str, Any] = {}
        unknown: List[str] = []
        i = 0
        argv = list(argv)
        while i < len(argv):
            tok = argv[i]
            if not tok.startswith("-") or len(tok) < 2:
                raise FlagError(f"expected a -flag but got {tok!r}")
            name = tok.lstrip("-")
            if i + 1 >= len(argv):
                raise FlagError(f"flag -{name} has no value")
            raw = argv[i + 1]
            i += 2
            p = self._params.get(name)
            if p is None:
                if not allow_unknown:
                    raise FlagError(f"unknown flag -{name}")
                unknown.extend([tok, raw])
                continue
            try:
                values[name] = p.parse(raw)
            except Exception as e:  # noqa: BLE001
                raise FlagError(f"bad value for -{name}: {raw!r} ({e})") from e
        if apply_defaults:
            for p in self._params.values():
                if p.name not in values:
                    if p.required:
                        raise FlagError(f"required flag -{p.name} is missing")
                    if p.default is not None:
                        values[p.name] = p.default
        return Configuration(values), unknown

    def usage(self) -> str:
        lines = []
        for p in sorted(self._params.values(), key=lambda q: (q.group, q.name)):
            lines.append(f"  -{p.name:<36} [{p.group}] {p.type.__name__} (default {p.default}) {p.doc}")
        return "\n".join(lines)


def config_to_argv(conf: Mapping[str, Any]) -> List[str]:
    out: List[str] = []
    for k, v in conf.items():
        if v is None:
            continue
        out.extend([f"-{k}", str(v).lower() if isinstance(v, bool) else str(v)])
    return out
