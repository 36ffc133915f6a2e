"""Run-time protobuf / gRPC stubs for the server's wire protocol.

The reference keeps `proto/stt.proto` in the tree but not the generated modules: `gen/stt/python/v1/__init__.py:3-9`
imports `stt_pb2` / `stt_pb2_grpc`, which its build makes with `grpc_tools.protoc`.  Where `grpc_tools` is not installed
(this image) nothing under `stt_server.backend.*` imports.  `install(proto_path)` builds the same two modules from the
`.proto` text with the protobuf runtime alone -- a small proto3 parser -> `FileDescriptorProto` -> message classes, plus
the servicer / stub / registration functions in the shape `protoc --grpc_python_out` emits -- and registers them under
the names the reference imports, so the unmodified server (and its load-test client) can run next to this backend.

Supported proto3 subset: package, enums, (nested) messages, scalar / enum / message fields, `optional`, `repeated`,
`map<k, v>`, `oneof`, `reserved`, services with unary and streaming rpcs.  No imports, options or extensions.
"""
from __future__ import annotations

import re
import sys
import types
from typing import Dict, List, Optional, Tuple

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
SCALARS = {
    "double": _F.TYPE_DOUBLE, "float": _F.TYPE_FLOAT, "int64": _F.TYPE_INT64, "uint64": _F.TYPE_UINT64,
    "int32": _F.TYPE_INT32, "fixed64": _F.TYPE_FIXED64, "fixed32": _F.TYPE_FIXED32, "bool": _F.TYPE_BOOL,
    "string": _F.TYPE_STRING, "bytes": _F.TYPE_BYTES, "uint32": _F.TYPE_UINT32, "sfixed32": _F.TYPE_SFIXED32,
    "sfixed64": _F.TYPE_SFIXED64, "sint32": _F.TYPE_SINT32, "sint64": _F.TYPE_SINT64,
}
_TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|[A-Za-z_][\w.]*|-?\d+|[{}()<>=;,\[\]]')


def _tokens(text: str) -> List[str]:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return _TOKEN.findall(text)


class _Parser:
    def __init__(self, text: str, file_name: str):
        self.t = _tokens(text)
        self.i = 0
        self.fd = descriptor_pb2.FileDescriptorProto(name=file_name, syntax="proto3")
        self.package = ""
        # filled while parsing, finished in parse_proto(): fields whose type is a message / enum name, and proto3 `optional`
        # fields (each gets a synthetic one-field oneof placed after the real ones, as protoc does)
        self.unresolved: List[descriptor_pb2.FieldDescriptorProto] = []
        self.synthetic: List[Tuple[descriptor_pb2.DescriptorProto, descriptor_pb2.FieldDescriptorProto, str]] = []

    def peek(self) -> Optional[str]:
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self, want: Optional[str] = None) -> str:
        if self.i >= len(self.t):
            raise ValueError("unexpected end of .proto")
        tok = self.t[self.i]
        self.i += 1
        if want is not None and tok != want:
            raise ValueError(f".proto: expected {want!r}, got {tok!r}")
        return tok

    def skip_statement(self) -> None:
        depth = 0
        while True:
            tok = self.take()
            if tok == "{":
                depth += 1
            elif tok == "}":
                depth -= 1
                if depth == 0:
                    return
            elif tok == ";" and depth == 0:
                return

    def parse(self) -> descriptor_pb2.FileDescriptorProto:
        while self.peek() is not None:
            tok = self.take()
            if tok == "syntax":
                self.take("=")
                if self.take().strip('"') != "proto3":
                    raise ValueError("only proto3 is supported")
                self.take(";")
            elif tok == "package":
                self.package = self.take()
                self.fd.package = self.package
                self.take(";")
            elif tok == "message":
                self.message(self.fd.message_type.add(), self.package)
            elif tok == "enum":
                self.enum(self.fd.enum_type.add())
            elif tok == "service":
                self.service()
            elif tok in ("option", "import"):
                if tok == "import":
                    raise ValueError(".proto imports are not supported")
                self.skip_statement()
            elif tok == ";":
                continue
            else:
                raise ValueError(f".proto: unexpected token {tok!r}")
        return self.fd

    def enum(self, ed: descriptor_pb2.EnumDescriptorProto) -> None:
        ed.name = self.take()
        self.take("{")
        while self.peek() != "}":
            name = self.take()
            if name in ("option", "reserved"):
                self.i -= 1
                self.skip_statement()
                continue
            self.take("=")
            ed.value.add(name=name, number=int(self.take()))
            if self.peek() == "[":
                while self.take() != "]":
                    pass
            self.take(";")
        self.take("}")

    def message(self, md: descriptor_pb2.DescriptorProto, scope: str) -> None:
        md.name = self.take()
        full = f"{scope}.{md.name}" if scope else md.name
        self.take("{")
        while self.peek() != "}":
            tok = self.take()
            if tok == "message":
                self.message(md.nested_type.add(), full)
            elif tok == "enum":
                self.enum(md.enum_type.add())
            elif tok in ("option", "reserved", "extensions"):
                self.i -= 1
                self.skip_statement()
            elif tok == "oneof":
                idx = len(md.oneof_decl)
                md.oneof_decl.add(name=self.take())
                self.take("{")
                while self.peek() != "}":
                    self.field(md, full, self.take(), oneof_index=idx)
                self.take("}")
            elif tok == ";":
                continue
            else:
                self.field(md, full, tok)
        self.take("}")

    def field(self, md: descriptor_pb2.DescriptorProto, full: str, first: str, oneof_index: Optional[int] = None) -> None:
        label, proto3_optional = _F.LABEL_OPTIONAL, False
        if first == "repeated":
            label, first = _F.LABEL_REPEATED, self.take()
        elif first == "optional":
            proto3_optional, first = True, self.take()
        if first == "map":
            self.take("<")
            key_t = self.take()
            self.take(",")
            val_t = self.take()
            self.take(">")
            name = self.take()
            self.take("=")
            number = int(self.take())
            entry = md.nested_type.add(name="".join(p.capitalize() for p in name.split("_")) + "Entry")
            entry.options.map_entry = True
            self._typed(entry.field.add(name="key", number=1, label=_F.LABEL_OPTIONAL, json_name="key"), key_t)
            self._typed(entry.field.add(name="value", number=2, label=_F.LABEL_OPTIONAL, json_name="value"), val_t)
            md.field.add(name=name, number=number, label=_F.LABEL_REPEATED, type=_F.TYPE_MESSAGE,
                         type_name=f".{full}.{entry.name}", json_name=_json_name(name))
        else:
            name = self.take()
            self.take("=")
            number = int(self.take())
            fd = md.field.add(name=name, number=number, label=label, json_name=_json_name(name))
            self._typed(fd, first)
            if oneof_index is not None:
                fd.oneof_index = oneof_index
            if proto3_optional:  # proto3 `optional` = a synthetic one-field oneof (what protoc emits)
                fd.proto3_optional = True
                self.synthetic.append((md, fd, name))
        if self.peek() == "[":
            while self.take() != "]":
                pass
        self.take(";")

    def _typed(self, fd: descriptor_pb2.FieldDescriptorProto, type_name: str) -> None:
        if type_name in SCALARS:
            fd.type = SCALARS[type_name]
        else:
            # resolved after parsing (enum vs message); proto3 files in scope here use package-level names
            fd.type_name = type_name
            self.unresolved.append(fd)

    def service(self) -> None:
        sd = self.fd.service.add(name=self.take())
        self.take("{")
        while self.peek() != "}":
            tok = self.take()
            if tok == "option":
                self.i -= 1
                self.skip_statement()
                continue
            if tok != "rpc":
                raise ValueError(f".proto: unexpected token {tok!r} in service")
            m = sd.method.add(name=self.take())
            self.take("(")
            if self.peek() == "stream":
                self.take()
                m.client_streaming = True
            m.input_type = self._qualify(self.take())
            self.take(")")
            self.take("returns")
            self.take("(")
            if self.peek() == "stream":
                self.take()
                m.server_streaming = True
            m.output_type = self._qualify(self.take())
            self.take(")")
            if self.peek() == "{":
                self.skip_statement()
            else:
                self.take(";")
        self.take("}")

    def _qualify(self, name: str) -> str:
        if name.startswith("."):
            return name
        return f".{self.package}.{name}" if self.package and "." not in name else f".{name}"


def _json_name(name: str) -> str:
    parts = name.split("_")
    return parts[0] + "".join(p.capitalize() for p in parts[1:])


def parse_proto(text: str, file_name: str = "stt.proto") -> descriptor_pb2.FileDescriptorProto:
    """proto3 text -> FileDescriptorProto (the subset described in the module docstring)."""
    p = _Parser(text, file_name)
    fd = p.parse()
    # type names: look them up among the declared messages / enums (outermost scope first, then nested)
    enums: Dict[str, str] = {}
    messages: Dict[str, str] = {}

    def walk(md: descriptor_pb2.DescriptorProto, scope: str) -> None:
        full = f"{scope}.{md.name}"
        messages[full] = full
        for e in md.enum_type:
            enums[f"{full}.{e.name}"] = f"{full}.{e.name}"
        for n in md.nested_type:
            walk(n, full)

    root = f".{fd.package}" if fd.package else ""
    for e in fd.enum_type:
        enums[f"{root}.{e.name}"] = f"{root}.{e.name}"
    for m in fd.message_type:
        walk(m, root)

    def resolve(name: str) -> Tuple[str, int]:
        cands = [name] if name.startswith(".") else [f"{root}.{name}"] + [k for k in list(messages) + list(enums) if k.endswith("." + name)]
        for c in cands:
            if c in messages:
                return c, _F.TYPE_MESSAGE
            if c in enums:
                return c, _F.TYPE_ENUM
        raise ValueError(f".proto: unknown type {name!r}")

    for f in p.unresolved:
        f.type_name, f.type = resolve(f.type_name)
    for md, f, name in p.synthetic:  # synthetic oneofs go after the real ones
        f.oneof_index = len(md.oneof_decl)
        md.oneof_decl.add(name=f"_{name}")
    return fd


def build_pb2_module(fd: descriptor_pb2.FileDescriptorProto, module_name: str) -> types.ModuleType:
    """Message classes, enum values and DESCRIPTOR under the names a protoc-generated `*_pb2` module exposes."""
    pool = descriptor_pool.DescriptorPool()
    file_desc = pool.Add(fd) if hasattr(pool, "Add") else None
    if file_desc is None:
        file_desc = pool.FindFileByName(fd.name)
    mod = types.ModuleType(module_name)
    mod.DESCRIPTOR = file_desc
    for name, desc in file_desc.message_types_by_name.items():
        setattr(mod, name, message_factory.GetMessageClass(desc))
    for name, enum in file_desc.enum_types_by_name.items():
        from google.protobuf.internal import enum_type_wrapper

        setattr(mod, name, enum_type_wrapper.EnumTypeWrapper(enum))
        for v in enum.values:
            setattr(mod, v.name, v.number)
    return mod


def build_grpc_module(pb2: types.ModuleType, module_name: str) -> types.ModuleType:
    """`<Service>Stub`, `<Service>Servicer`, `add_<Service>Servicer_to_server` for every service of the file."""
    import grpc

    mod = types.ModuleType(module_name)
    for svc in pb2.DESCRIPTOR.services_by_name.values():
        methods = []
        for m in svc.methods:
            req = getattr(pb2, m.input_type.name)
            rsp = getattr(pb2, m.output_type.name)
            kind = ("stream" if m.client_streaming else "unary") + "_" + ("stream" if m.server_streaming else "unary")
            methods.append((m.name, f"/{svc.full_name}/{m.name}", kind, req, rsp))

        def stub_init(self, channel, _methods=methods):
            for name, path, kind, req, rsp in _methods:
                setattr(self, name, getattr(channel, kind)(path, request_serializer=req.SerializeToString,
                                                           response_deserializer=rsp.FromString))

        def make_unimplemented(name):
            def method(self, request, context):
                context.set_code(grpc.StatusCode.UNIMPLEMENTED)
                context.set_details("Method not implemented!")
                raise NotImplementedError("Method not implemented!")

            method.__name__ = name
            return method

        servicer = type(f"{svc.name}Servicer", (object,), {name: make_unimplemented(name) for name, *_ in methods})

        def add_to_server(servicer_obj, server, _methods=methods, _svc=svc.full_name):
            handlers = {}
            for name, _path, kind, req, rsp in _methods:
                factory = getattr(grpc, f"{kind}_rpc_method_handler")
                handlers[name] = factory(getattr(servicer_obj, name), request_deserializer=req.FromString,
                                         response_serializer=rsp.SerializeToString)
            server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(_svc, handlers),))

        setattr(mod, f"{svc.name}Stub", type(f"{svc.name}Stub", (object,), {"__init__": stub_init}))
        setattr(mod, f"{svc.name}Servicer", servicer)
        setattr(mod, f"add_{svc.name}Servicer_to_server", add_to_server)
    return mod


def install(proto_path: str, package: str = "gen.stt.python.v1") -> Tuple[types.ModuleType, types.ModuleType]:
    """Build `stt_pb2` / `stt_pb2_grpc` from `proto_path` and register them as `<package>.stt_pb2[_grpc]` (and under the
    bare names, as gen/stt/python/v1/__init__.py:5-9 does) unless real generated modules are importable already."""
    try:
        __import__(f"{package}.stt_pb2")
        return sys.modules[f"{package}.stt_pb2"], sys.modules[f"{package}.stt_pb2_grpc"]
    except Exception:  # noqa: BLE001 - absent or half-initialised package: build the modules ourselves
        for name in [n for n in sys.modules if n == package or n.startswith(package + ".")]:
            sys.modules.pop(name, None)
    with open(proto_path, "r", encoding="utf-8") as fh:
        fd = parse_proto(fh.read(), proto_path.rsplit("/", 1)[-1])
    pb2 = build_pb2_module(fd, f"{package}.stt_pb2")
    pb2_grpc = build_grpc_module(pb2, f"{package}.stt_pb2_grpc")
    parts = package.split(".")
    for depth in range(1, len(parts) + 1):  # parent packages as plain namespace modules
        name = ".".join(parts[:depth])
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = []  # type: ignore[attr-defined]
            sys.modules[name] = pkg
        if depth > 1:
            setattr(sys.modules[".".join(parts[: depth - 1])], parts[depth - 1], sys.modules[name])
    leaf = sys.modules[package]
    leaf.stt_pb2, leaf.stt_pb2_grpc = pb2, pb2_grpc
    sys.modules[f"{package}.stt_pb2"] = pb2
    sys.modules[f"{package}.stt_pb2_grpc"] = pb2_grpc
    sys.modules.setdefault("stt_pb2", pb2)
    sys.modules.setdefault("stt_pb2_grpc", pb2_grpc)
    return pb2, pb2_grpc
