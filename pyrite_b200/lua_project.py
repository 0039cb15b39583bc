"""Load a pyrite `project.lua` unmodified.

pyrite's `load_project` (pyrite/src/project/mod.rs:29-93) puts the project directory on `package.path`,
registers `assign_id`, runs its DSL library `lib.lua` and then evaluates the project file; the returned
table is decoded into the typed `Project`.  Here the project file runs in pyrite_b200.lua's interpreter
and the DSL library is provided natively: the same globals with the same behaviour (`shape.*`,
`material.*`, `ray_marched.*`, `quaternion_julia.cubic`, `bounds.box`, `transform.look_at`,
`camera.perspective`, `renderer.*`, `light.point`, `light_source.d65 / .a`, `vector`, `rgb`, `spectrum`,
`texture`, `blackbody`, `fresnel`, `mix`, the arithmetic metamethods on expressions, `:clone()` /
`:with{}` on every DSL object, `dump`, `assign_id`, `_pyrite.*`).  The resulting table is converted to
the project table `pyrite_b200.project.serialize_project` decodes, preserving table identity (node
interning by identity is what `typed_nodes` does through `assign_id`).
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict

from . import project as P
from .lua import Interpreter, LuaError, LuaFunction, LuaTable, lua_type, tostring


class ProjectLoadError(ValueError):
    pass


def _install_dsl(L: Interpreter):
    G = L.G
    ids = {"next": 0}

    def assign_id(t):
        # project/tables.rs:14-18: every DSL object carries a unique `_id`
        ids["next"] += 1
        t.set("_id", ids["next"])

    basics_mt = LuaTable()
    basics_mt.set("__index", basics_mt)
    expression_mt = LuaTable()
    expression_mt.set("__index", expression_mt)
    expression_fallback = LuaTable()
    expression_fallback.set("__index", basics_mt)
    expression_mt.meta = expression_fallback

    def make_object(obj, meta):
        obj.meta = meta
        assign_id(obj)

    def make_basic(obj):
        make_object(obj, basics_mt)

    def make_expression(obj):
        make_object(obj, expression_mt)

    def clone(self):
        if not isinstance(self, LuaTable):
            return [self]
        c = LuaTable()
        for k, v in self.items():
            c.set(k, v)
        make_object(c, self.meta)
        return [c]

    def with_(self, new_properties=None):
        c = clone(self)[0]
        props = L.call(new_properties, [c])[0] if callable(new_properties) else new_properties
        if not isinstance(props, LuaTable):
            raise LuaError(f"bad argument #1 to 'with' (table expected, got {lua_type(props)})")
        for k, v in props.items():
            c.set(k, v)
        return [c]

    basics_mt.set("clone", clone)
    basics_mt.set("with", with_)

    def binary(operator):
        def op(lhs, rhs):
            t = LuaTable()
            t.set("type", "binary"); t.set("operator", operator); t.set("lhs", lhs); t.set("rhs", rhs)
            make_expression(t)
            return [t]

        return op

    for event, name in (("__add", "add"), ("__sub", "sub"), ("__mul", "mul"), ("__div", "div")):
        expression_mt.set(event, binary(name))

    def expression(type_name, **fields):
        t = LuaTable()
        t.set("type", type_name)
        for k, v in fields.items():
            t.set(k, v)
        make_expression(t)
        return t

    def mix(self, other=None, amount=None):
        if isinstance(self, LuaTable) and self.get("type") is None:
            self.set("type", "mix")
            make_expression(self)
            return [self]
        return [expression("mix", lhs=self, rhs=other, amount=amount)]

    expression_mt.set("mix", mix)
    G.set("mix", mix)
    G.set("fresnel", lambda ior=None, env_ior=None: [expression("fresnel", ior=ior, env_ior=env_ior if env_ior is not None and env_ior is not False else 1)])

    def or0(v):
        return v if (v is not None and v is not False) else 0.0

    def vector(x=None, y=None, z=None, w=None):
        if isinstance(x, LuaTable) and x.get("type") is None:
            return [expression("vector", x=or0(x.get("x")), y=or0(x.get("y")), z=or0(x.get("z")), w=or0(x.get("w")))]
        return [expression("vector", x=or0(x), y=or0(y), z=or0(z), w=or0(w))]

    G.set("vector", vector)
    G.set("blackbody", lambda temperature=None: [expression("blackbody", temperature=temperature)])
    G.set("rgb", lambda r=None, g=None, b=None: [expression("rgb", red=or0(r), green=or0(g), blue=or0(b))])

    def tagged(type_name, maker):
        def make(properties=None):
            if not isinstance(properties, LuaTable):
                raise LuaError(f"attempt to index a {lua_type(properties)} value (local 'properties')")
            properties.set("type", type_name)
            maker(properties)
            return [properties]

        return make

    G.set("spectrum", tagged("spectrum", make_expression))

    def texture(path=None, *modifiers):
        t = expression("color_texture", path=path, linear=False, mono=False)
        for m in modifiers:
            if isinstance(m, str):
                t.set(m, True)
        if t.get("mono") is True:
            t.set("type", "mono_texture")
        return [t]

    G.set("texture", texture)

    def namespace(**entries):
        t = LuaTable()
        for k, v in entries.items():
            t.set(k, v)
        return t

    G.set("shape", namespace(sphere=tagged("sphere", make_basic), plane=tagged("plane", make_basic), mesh=tagged("mesh", make_basic),
                             ray_marched=tagged("ray_marched", make_basic)))
    G.set("ray_marched", namespace(quaternion_julia=tagged("quaternion_julia", make_basic), mandelbulb=tagged("mandelbulb", make_basic)))
    cubic = LuaTable()
    cubic.set("type", "quaternion_julia"); cubic.set("name", "cubic")
    make_basic(cubic)
    G.set("quaternion_julia", namespace(cubic=cubic))
    G.set("bounds", namespace(box=tagged("box", make_basic)))
    G.set("material", namespace(diffuse=tagged("diffuse", make_expression), emissive=tagged("emissive", make_expression),
                                mirror=tagged("mirror", make_expression), refractive=tagged("refractive", make_expression)))
    G.set("light_source", namespace(d65=expression("spectrum", name="d65"), a=expression("spectrum", name="a")))
    G.set("transform", namespace(look_at=tagged("look_at", make_basic)))
    G.set("camera", namespace(perspective=tagged("perspective", make_basic)))
    G.set("renderer", namespace(simple=tagged("simple", make_basic), bidirectional=tagged("bidirectional", make_basic),
                                photon_mapping=tagged("photon_mapping", make_basic)))
    G.set("light", namespace(point=tagged("point_light", make_basic)))
    G.set("assign_id", lambda t: assign_id(t))

    def dump(o=None, tabs=None):
        tabs = int(tabs) if tabs else 1
        if not isinstance(o, LuaTable):
            return [tostring(o)]
        s = "{\n"
        for k, v in o.items():
            key = tostring(k) if isinstance(k, (int, float)) and not isinstance(k, bool) else '"' + tostring(k) + '"'
            s += "  " * tabs + "[" + key + "] = " + dump(v, tabs + 1)[0] + ",\n"
        return [s + "  " * (tabs - 1) + "}"]

    G.set("dump", dump)
    internals = namespace(basics_mt=basics_mt, expression_mt=expression_mt, make_basic=lambda o: make_basic(o),
                          make_expression=lambda o: make_expression(o), make_object=lambda o, m: make_object(o, m))
    internals.set("binary_operator", lambda op, lhs, rhs: binary(op)(lhs, rhs))
    G.set("_pyrite", internals)


def _to_python(value, seen: Dict[int, Any], as_expr_hint: bool = False):
    """LuaTable tree -> the Node / Expr / list structure pyrite_b200.project.serialize_project decodes.
    One Python object per Lua table (identity is what interns expression and material nodes)."""
    if not isinstance(value, LuaTable):
        if isinstance(value, LuaFunction) or callable(value):
            raise ProjectLoadError("functions cannot be part of a project description")
        return value
    if id(value) in seen:
        return seen[id(value)]
    n = value.length()
    keys = list(value.hash.keys())
    is_sequence = n > 0 and len(keys) == n
    if is_sequence:
        out_list: list = []
        seen[id(value)] = out_list
        for i in range(1, n + 1):
            out_list.append(_to_python(value.hash[i], seen))
        return out_list
    is_expression = value.meta is not None and "__add" in value.meta.hash
    node = P.Expr() if is_expression else P.Node()
    seen[id(value)] = node
    for k, v in value.hash.items():
        if k == "_id":
            continue
        node[k if isinstance(k, str) else tostring(k)] = _to_python(v, seen)
    return node


def _install_host_functions(L: Interpreter):
    """What the Rust side registers before it runs its DSL library (project/tables.rs:14-18): `assign_id` only."""
    ids = {"next": 0}

    def assign_id(t):
        ids["next"] += 1
        t.set("_id", ids["next"])

    L.G.set("assign_id", lambda t: assign_id(t))


def load_project(path, output=print, dsl_library=None) -> tuple:
    """Evaluate `path` (a pyrite project.lua) and return (project_table, project_dir).

    `dsl_library`: path of a Lua DSL library to run verbatim before the project file instead of the native DSL
    (pyrite's own `src/project/lib.lua`, as `load_project` does, project/mod.rs:43-60) - used by the differential test
    that checks the native DSL against it."""
    path = Path(path)
    project_dir = path.resolve().parent
    L = Interpreter(search_dirs=[project_dir], output=output)
    try:
        if dsl_library is None:
            _install_dsl(L)
        else:
            _install_host_functions(L)
            L.current_chunk = Path(dsl_library).name
            L.run(Path(dsl_library).read_text(), Path(dsl_library).name)
    except LuaError as e:
        raise ProjectLoadError(f"error while running the DSL library: {e}") from e
    L.current_chunk = path.name
    try:
        result = L.run(path.read_text(), path.name)
    except LuaError as e:
        raise ProjectLoadError(f"error while running {path.name}: {e}") from e
    except RecursionError as e:
        raise ProjectLoadError(f"error while running {path.name}: stack overflow") from e
    if not result or not isinstance(result[0], LuaTable):
        raise ProjectLoadError(f"{path.name} did not return a project table")
    table = _to_python(result[0], {})
    if isinstance(table, list):
        raise ProjectLoadError(f"{path.name} returned a sequence, not a project table")
    return table, project_dir


def load_project_ir(path, output=print, dsl_library=None) -> bytes:
    """project.lua -> project IR blob for pyr_project_load."""
    table, project_dir = load_project(path, output, dsl_library)
    try:
        return P.serialize_project(table, base_dir=project_dir)
    except P.ProjectError as e:
        raise ProjectLoadError(str(e)) from e
