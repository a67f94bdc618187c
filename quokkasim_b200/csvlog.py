"""Binary log records of one replication -> the reference's CSV files.

`logger.write_csv(dir)` in the reference drains the logger's EventQueue through `csv::Writer::serialize`
(core.rs:388-401).  Here the records of every component connected to the logger are decoded from the 32-byte
`qk_log_record` stream that the step kernel wrote, and formatted field for field like the reference's hand-written
`Serialize` impls:

  DiscreteStockLog     discrete.rs:295-318  time,event_id,source_event_id,element_name,element_type,log_type,item,reason
  DiscreteProcessLog   discrete.rs:608-633  time,event_id,source_event_id,element_name,element_type,event_type,item,reason
  VectorStockLog       vector.rs:233-256    ...,log_type,balance,vector_total,vector,new_state
  VectorProcessLog     vector.rs:594-720    ...,event_type,total,inflows,outflows,reason
  BasicEnvironmentLog  core.rs:306-325      time,event_id,source_event_id,element_name,element_type,event

The container family logs through DiscreteStockLog<Vector3Container> / DiscreteProcessLog<Vector3Container>: same
columns, the `item` column is `serde_json::to_string` of the container (Serialize impl vector_container.rs:113-124):
`{"id":"Truck_01","resource":{"x":1.0,"y":2.0,"z":3.0},"capacity":80.0}` with `"resource":null` for None.

Formatting facts this relies on (third-party crates, not in /root/reference -- restated from their documented
behaviour): the `csv` crate quotes a field only when it contains `,` `"` CR or LF and doubles embedded quotes,
terminates records with LF, writes `None` as an empty field; `serde_json` and `csv` print f64 with `ryu`
(shortest round-trip digits, `1.0` not `1`, exponent form `1e21` / `1.5e-7` outside [1e-5, 1e21)); chrono's
`DateTime<Utc>` Display is `%Y-%m-%d %H:%M:%S%.f UTC` with 0/3/6/9 fractional digits; tai-time's `MonotonicTime`
Display (used by BasicEnvironment only, core.rs:277) is `%Y-%m-%d %H:%M:%S` plus `.fffffffff` when the nanosecond
part is not zero.

Rows are written in canonical schedule order; the reference's file order is the order in which
its concurrently running models pushed into the shared EventQueue, which equals this order on a single-threaded
executor.
"""
import datetime
import os
import struct

from . import _capi as capi

_REASONS = {
    1: "Upstream is empty", 2: "Upstream is not connected", 3: "Downstream is full", 4: "Downstream is not connected",
    5: "Upstream did not provide resource", 6: "Stopped by environment", 7: "Resumed by environment",
    8: "Upstream is full", 9: "At least one upstream is empty", 10: "No upstreams are connected",
    11: "At least one downstream is full", 12: "No downstreams are connected",
    13: "No remaining process slots", 14: "Downstream container stock is full",
    15: "Downstream resource stock is full",
}
_STATE = ("Empty", "Normal", "Full")
_EPOCH = datetime.datetime(1970, 1, 1)

K = capi  # log kinds live in _capi as LOG_*


def ryu_f64(x):
    """f64 -> text the way `ryu::Buffer::format` prints it (used by serde_json and the csv crate)."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    if x == 0.0:
        return "-0.0" if struct.pack("<d", x)[7] & 0x80 else "0.0"
    sign = "-" if x < 0 else ""
    r = repr(abs(x))
    if "e" in r:
        mant, exp = r.split("e")
        exp = int(exp)
    else:
        mant, exp = r, 0
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first digit: value = 0.d1d2... * 10^kk
    kk = len(ip.lstrip("0")) + exp if ip.strip("0") else exp - (len(fp) - len(fp.lstrip("0")))
    digits = digits.rstrip("0") or "0"
    n = len(digits)
    if 0 < kk <= 16 and n <= kk:  # 1234e7 -> 12340000000.0
        return sign + digits + "0" * (kk - n) + ".0"
    if 0 < kk <= 16:  # 1234e-2 -> 12.34
        return sign + digits[:kk] + "." + digits[kk:]
    if -5 < kk <= 0:  # 1234e-6 -> 0.001234
        return sign + "0." + "0" * (-kk) + digits
    e = kk - 1
    if n == 1:
        return "%s%se%d" % (sign, digits, e)
    return "%s%s.%se%d" % (sign, digits[0], digits[1:], e)


def chrono_utc(t_ns):
    """MonotonicTime::to_chrono_date_time(0).unwrap().to_string()  (discrete.rs:239, vector.rs:177)"""
    secs, nanos = divmod(int(t_ns), 1_000_000_000)
    d = _EPOCH + datetime.timedelta(seconds=secs)
    s = d.strftime("%Y-%m-%d %H:%M:%S")
    if nanos:
        if nanos % 1_000_000 == 0:
            s += ".%03d" % (nanos // 1_000_000)
        elif nanos % 1_000 == 0:
            s += ".%06d" % (nanos // 1_000)
        else:
            s += ".%09d" % nanos
    return s + " UTC"


def tai_display(t_ns):
    """MonotonicTime's own Display (core.rs:277 `now.to_string()`)."""
    secs, nanos = divmod(int(t_ns), 1_000_000_000)
    s = (_EPOCH + datetime.timedelta(seconds=secs)).strftime("%Y-%m-%d %H:%M:%S")
    return s + (".%09d" % nanos if nanos else "")


def csv_field(v):
    if v is None:
        return ""
    s = str(v)
    if any(ch in s for ch in ',"\r\n'):
        return '"' + s.replace('"', '""') + '"'
    return s


def json_string(s):
    """serde_json::to_string(&String)"""
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif ch == "\b":
            out.append("\\b")
        elif ch == "\f":
            out.append("\\f")
        elif o < 0x20:
            out.append("\\u%04x" % o)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def _json_f64(x):
    """serde_json prints non-finite floats as null"""
    x = float(x)
    return ryu_f64(x) if x == x and x not in (float("inf"), float("-inf")) else "null"


class LogDecoder:
    """Turns the records of one replication into reference-format rows, per logger."""

    def __init__(self, sim):
        self.sim = sim
        self.components = sim.components
        self.variants = sim.variants
        sources = [c for c in self.components if c.KIND == capi.DISCRETE_SOURCE]
        self.sources = sources
        self.initial_items = []
        for c in self.components:
            if c.KIND == capi.DISCRETE_STOCK:
                self.initial_items.extend(list(c.resource))

    # -- pieces -----------------------------------------------------------------------------------------
    def event_id(self, comp, seq):
        if comp == 0xFFFF:
            return "INIT_000000"
        if comp == 0xFFFE:
            return "SCH_000000"
        return "%s_%06d" % (self.components[comp].element_code, seq)

    def item_name(self, handle):
        handle = int(handle)
        ordn, idx = (handle >> 26) & 63, handle & 0x3FFFFFF
        if ordn == 63:
            return str(self.initial_items[idx])
        f = self.sources[ordn].item_factory
        return f.item_name(idx)

    def is_container(self, comp):
        return self.variants[comp].startswith("Vector3Container")

    def is_vector(self, comp):
        from . import vector_api

        return self.variants[comp] in vector_api.VECTOR_VARIANTS

    def item_json(self, comp, handle, rec_vecs):
        """serde_json::to_string(item): a String item, or a Vector3Container (vector_container.rs:113-124)"""
        if not self.is_container(comp):
            return json_string(self.item_name(handle))
        handle = int(handle)
        ordn, idx = (handle >> 26) & 63, handle & 0x3FFFFFF
        if ordn == 63:
            it = self.initial_items[idx]
            name, capacity = it.id, it.capacity
        else:
            f = self.sources[ordn].item_factory
            name, capacity = f.item_name(idx), f.capacity
        raw = rec_vecs[0] if rec_vecs else None
        if raw is None or raw[0] == capi.CONTAINER_NONE_BITS:
            res = "null"
        else:
            res = '{"x":%s,"y":%s,"z":%s}' % tuple(_json_f64(self._f64(w)) for w in raw)
        return '{"id":%s,"resource":%s,"capacity":%s}' % (json_string(name), res, _json_f64(capacity))

    def vector_json(self, comp, v3):
        c = self.components[comp]
        if c.dim == 1:
            return ryu_f64(v3[0])
        rt = getattr(c, "resource_type", None)
        if rt is not None:  # #[derive(Serialize)] on the user's struct: its fields by name, in declaration order
            return "{" + ",".join('"%s":%s' % (f, ryu_f64(x)) for f, x in zip(rt.FIELDS, v3)) + "}"
        return '{"x":%s,"y":%s,"z":%s}' % tuple(ryu_f64(x) for x in v3[:3])

    def vector_total(self, comp, v3):
        c = self.components[comp]
        if c.dim == 1:
            return v3[0]
        rt = getattr(c, "resource_type", None)
        if rt is not None:
            return rt(*v3[: c.dim]).total()
        return (v3[0] + v3[1]) + v3[2]

    @staticmethod
    def _f64(bits):
        return struct.unpack("<d", struct.pack("<Q", int(bits)))[0]

    @staticmethod
    def _vec_of(rec):
        raw = rec.tobytes()
        return (struct.unpack_from("<d", raw, 0)[0], struct.unpack_from("<d", raw, 16)[0],
                struct.unpack_from("<d", raw, 24)[0])

    # -- rows -------------------------------------------------------------------------------------------
    def _grouped(self, records):
        """yield (main record, kind, vectors, raw words): every main record with its continuation records folded in."""
        i, n = 0, len(records)
        while i < n:
            r = records[i]
            i += 1
            kind = int(r["kind"])
            if kind == capi.LOG_VEC_DATA:
                continue  # orphan (its main record fell out of the ring)
            vecs, raws = [], []
            while i < n and int(records[i]["kind"]) == capi.LOG_VEC_DATA and int(records[i]["comp"]) == int(r["comp"]) \
                    and int(records[i]["seq"]) == int(r["seq"]):
                chunk = int(records[i]["reason"]) >> 3  # resources of more than three fields: index + 8 * chunk
                if chunk == 0:
                    vecs.append(self._vec_of(records[i]))
                elif vecs:
                    vecs[int(records[i]["reason"]) & 7] += self._vec_of(records[i])
                q = struct.unpack("<QQQQ", records[i].tobytes())
                if chunk == 0:
                    raws.append((q[0], q[2], q[3]))  # the three fp64 words as raw bits (container None sentinel)
                i += 1
            yield r, kind, vecs, raws

    def rows(self, records):
        """yield (component id, [fields...]) for every main record, continuation records folded in."""
        for r, kind, vecs, raws in self._grouped(records):
            yield int(r["comp"]), self._row(r, kind, vecs, raws)

    def vector_events(self, records):
        """yield (component id, dict) for the records of the continuous components, typed like the reference's
        VectorStockLog<T> / VectorProcessLog<T> (vector.rs:190-209, 543-592): what a user's own log struct is built
        from when the logger is connected with `map_connect_sink(|c| c.clone().into(), ..)` (iron_ore.rs:143-155,
        316-324).  `vector` / `vectors` are instances of the component's resource type (float, Vector3 or the class
        made by define_resource)."""
        from . import vector_api

        for r, kind, vecs, _ in self._grouped(records):
            comp = int(r["comp"])
            if not self.is_vector(comp):
                continue
            c = self.components[comp]
            rt = getattr(c, "resource_type", None)

            def typed(v, c=c, rt=rt):
                return v[0] if c.dim == 1 else (rt(*v[: c.dim]) if rt is not None else vector_api.Vector3(v[:3]))

            ev = {"time": chrono_utc(int(r["time_ns"])), "event_id": self.event_id(comp, int(r["seq"])),
                  "source_event_id": self.event_id(int(r["src_comp"]), int(r["src_seq"])),
                  "element_name": c.element_name, "element_type": c.element_type}
            reason, scalar = int(r["reason"]), self._f64(int(r["payload"]))
            if kind in (capi.LOG_VSTOCK_ADD, capi.LOG_VSTOCK_REMOVE):
                ev.update(log_type="Add" if kind == capi.LOG_VSTOCK_ADD else "Remove", balance=scalar,
                          vector=typed(vecs[0]))
            elif kind == capi.LOG_VSTOCK_STATE_CHANGE:
                ev.update(log_type="StateChange", new_state=_STATE[reason])
            elif kind in (capi.LOG_VPROC_START, capi.LOG_VPROC_COMBINE_START, capi.LOG_VPROC_SPLIT_START,
                          capi.LOG_VPROC_SUCCESS, capi.LOG_VPROC_COMBINE_SUCCESS, capi.LOG_VPROC_SPLIT_SUCCESS):
                names = {capi.LOG_VPROC_START: "ProcessStart", capi.LOG_VPROC_SUCCESS: "ProcessSuccess",
                         capi.LOG_VPROC_COMBINE_START: "CombineStart", capi.LOG_VPROC_COMBINE_SUCCESS: "CombineSuccess",
                         capi.LOG_VPROC_SPLIT_START: "SplitStart", capi.LOG_VPROC_SPLIT_SUCCESS: "SplitSuccess"}
                ev.update(event_type=names[kind], quantity=scalar, vectors=[typed(v) for v in vecs])
            elif kind == capi.LOG_VPROC_FAILURE:
                ev.update(event_type="ProcessFailure", reason=_REASONS.get(reason))
            elif kind == capi.LOG_WITHDRAW_REQUEST:
                ev.update(event_type="WithdrawRequest")
            elif kind in (capi.LOG_DELAY_START, capi.LOG_DELAY_END):
                ev.update(event_type="DelayStart" if kind == capi.LOG_DELAY_START else "DelayEnd",
                          reason=c.delay_modes[int(r["payload"])].name)
            else:
                simple = {capi.LOG_PROCESS_CONTINUE: "ProcessContinue", capi.LOG_PROCESS_NONSTART: "ProcessNonStart",
                          capi.LOG_PROCESS_STOPPED: "ProcessStopped"}
                ev.update(event_type=simple.get(kind, "Unknown"), reason=_REASONS.get(reason))
            yield comp, ev

    def _row(self, r, kind, vecs, raws=()):
        comp = int(r["comp"])
        c = self.components[comp]
        variant = self.variants[comp]
        head = [None, self.event_id(comp, int(r["seq"])), self.event_id(int(r["src_comp"]), int(r["src_seq"])),
                c.element_name, c.element_type]
        t = int(r["time_ns"])
        reason = int(r["reason"])
        payload = int(r["payload"])
        if kind == capi.LOG_ENV_STATE:
            head[0] = tai_display(t)
            head[4] = "BasicEnvironment"
            return head + ["Stopped" if reason == 1 else "Normal"]
        head[0] = chrono_utc(t)
        if kind == capi.LOG_STOCK_ADD:
            return head + ["Add", self.item_json(comp, payload, raws), None]
        if kind == capi.LOG_STOCK_REMOVE:
            item = "null" if payload == 0xFFFFFFFFFFFFFFFF else self.item_json(comp, payload, raws)
            return head + ["Remove", item, None]
        if kind == capi.LOG_STOCK_STATE_CHANGE:
            st = '{"%s":{"occupied":%d,"empty":%d}}' % (_STATE[reason], payload >> 32, payload & 0xFFFFFFFF)
            return head + ["StateChange", st, None]
        if kind in (capi.LOG_PROCESS_START, capi.LOG_PROCESS_FINISH):
            name = "ProcessStart" if kind == capi.LOG_PROCESS_START else "ProcessFinish"
            return head + [name, self.item_json(comp, payload, raws), None]
        if kind in (capi.LOG_DELAY_START, capi.LOG_DELAY_END):
            name = "DelayStart" if kind == capi.LOG_DELAY_START else "DelayEnd"
            dm = c.delay_modes[payload].name
            if self.is_vector(comp):  # vector.rs:693-706: the name travels in `reason`
                return head + [name, None, None, None, dm]
            return head + [name, dm, None]
        simple = {capi.LOG_PROCESS_CONTINUE: "ProcessContinue", capi.LOG_PROCESS_NONSTART: "ProcessNonStart",
                  capi.LOG_PROCESS_STOPPED: "ProcessStopped", capi.LOG_WITHDRAW_REQUEST: "WithdrawRequest"}
        if kind in simple:
            why = _REASONS.get(reason) if kind != capi.LOG_WITHDRAW_REQUEST else None
            if self.is_vector(comp):
                return head + [simple[kind], None, None, None, why]
            return head + [simple[kind], None, why]
        # ---- vector stock
        if kind in (capi.LOG_VSTOCK_ADD, capi.LOG_VSTOCK_REMOVE):
            v = vecs[0] if vecs else (0.0, 0.0, 0.0)
            return head + ["add" if kind == capi.LOG_VSTOCK_ADD else "remove", ryu_f64(self._f64(payload)),
                           ryu_f64(self.vector_total(comp, v)), self.vector_json(comp, v), None]
        if kind == capi.LOG_VSTOCK_STATE_CHANGE:
            return head + ["emit_change", None, None, None, _STATE[reason]]
        # ---- vector process-like
        family = "Combine" if "Combiner" in variant else ("Split" if "Splitter" in variant else "Process")
        if kind == capi.LOG_VPROC_FAILURE:
            # combiners and splitters log VectorProcessLogType::ProcessFailure too (vector.rs:937-949, 1235-1247), which
            # serialises as "ProcessFailure" (vector.rs:621-622): only Start / Success carry the Combine / Split prefix
            return head + ["ProcessFailure", None, None, None, _REASONS.get(reason)]
        q = ryu_f64(self._f64(payload))
        js = "[" + ",".join(self.vector_json(comp, v) for v in vecs) + "]"
        if kind in (capi.LOG_VPROC_START, capi.LOG_VPROC_COMBINE_START, capi.LOG_VPROC_SPLIT_START):
            return head + [family + "Start", q, js, None, None]
        if kind in (capi.LOG_VPROC_SUCCESS, capi.LOG_VPROC_COMBINE_SUCCESS, capi.LOG_VPROC_SPLIT_SUCCESS):
            return head + [family + "Success", q, None, js, None]
        raise ValueError("unknown log record kind %d" % kind)


HEADERS = {
    "DiscreteStockLogger": "time,event_id,source_event_id,element_name,element_type,log_type,item,reason",
    "DiscreteProcessLogger": "time,event_id,source_event_id,element_name,element_type,event_type,item,reason",
    "VectorStockLogger": "time,event_id,source_event_id,element_name,element_type,log_type,balance,vector_total,"
                         "vector,new_state",
    "VectorProcessLogger": "time,event_id,source_event_id,element_name,element_type,event_type,total,inflows,"
                           "outflows,reason",
    "BasicEnvironmentLogger": "time,event_id,source_event_id,element_name,element_type,event",
}
HEADERS["ContainerStockLogger"] = HEADERS["DiscreteStockLogger"]
HEADERS["ContainerProcessLogger"] = HEADERS["DiscreteProcessLogger"]


def csv_text(component_logger, sim, replication=0, record_map=None, header=None):
    """The CSV file content `logger.write_csv(dir)` would produce for one replication.  `record_map` + `header`: a
    user-defined logger (iron_ore.rs:98-256) -- record_map(event) turns the typed event of vector_events() into the
    fields of the user's own log struct, `header` names them."""
    logger = component_logger.logger
    records, total = sim.read_log(replication)
    if total > len(records):
        raise RuntimeError("log ring kept %d of %d records of replication %d: raise log_capacity for CSV export"
                           % (len(records), total, replication))
    members = set(id(c) for c in logger.components)
    dec = LogDecoder(sim)
    lines = []
    if record_map is not None:
        for comp, ev in dec.vector_events(records):
            if id(sim.components[comp]) in members:
                lines.append(",".join(csv_field(f) for f in record_map(ev)))
    else:
        for comp, fields in dec.rows(records):
            if id(sim.components[comp]) in members:
                lines.append(",".join(csv_field(f) for f in fields))
    # csv::Writer with has_headers(true) writes the header with the first record only: an empty log is an empty file
    if not lines:
        return ""
    return (header if record_map is not None else HEADERS[type(logger).__name__]) + "\n" + "\n".join(lines) + "\n"


def write_csv(component_logger, directory, sim, replication=0, record_map=None, header=None):
    """Logger::write_csv (core.rs:388-401): File::create("{dir}/{name}.csv")."""
    path = os.path.join(directory, "%s.csv" % component_logger.logger.get_name())
    with open(path, "w", newline="") as f:
        f.write(csv_text(component_logger, sim, replication, record_map, header))
    return path
