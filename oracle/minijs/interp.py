"""TEST INFRASTRUCTURE ONLY.  Closure-compiling evaluator for the tuples of oracle/minijs/parser.py.

JS values:  Number -> Python float (always; IEEE double like JS), Boolean -> bool, String -> str, null -> None,
undefined -> UNDEF, Array -> JSArray, typed arrays -> JSTyped (stores round the way the element type says),
objects -> JSObject, functions -> JSFunction / Native, Promise -> JSPromise (then-callbacks run from an explicit
micro-task queue: Interp.run_microtasks(); window.setTimeout callbacks from Interp.run_timers()).
Math.log10 / Math.pow / `**` go through include/fa_jsmath.h (V8's fdlibm port), everything else is IEEE.
Runtime errors of the language (reading a property of undefined, calling a non-function) are thrown as JS
exceptions (JSThrow) so the reference's own try/catch and Promise-executor semantics see them."""
from __future__ import annotations

import math
import struct
from decimal import ROUND_HALF_UP, Decimal

from .. import jsmath
from .parser import Parser


class _Undefined:
    __slots__ = ()

    def __repr__(self):
        return 'undefined'

    def __bool__(self):
        return False


UNDEF = _Undefined()
NAN = float('nan')
INF = float('inf')


class JSThrow(Exception):
    def __init__(self, value):
        super().__init__(value)
        self.value = value

    def __str__(self):
        return 'JS exception: ' + to_str(self.value)


def type_error(msg):
    return JSThrow(JSObject({'name': 'TypeError', 'message': msg}))


class JSObject:
    __slots__ = ('props', 'getters')

    def __init__(self, props=None):
        self.props = props if props is not None else {}
        self.getters = None


class JSArray:
    __slots__ = ('a', 'props')

    def __init__(self, a=None):
        self.a = a if a is not None else []
        self.props = None         # named (non-index) properties, e.g. hist[NaN]++ creates "NaN"; for-in visits them


def _f32(x):
    try:
        return struct.unpack('f', struct.pack('f', x))[0]
    except OverflowError:
        return INF if x > 0 else -INF


def _i32(x):
    if x != x or x in (INF, -INF):
        return 0.0
    v = int(x) & 0xFFFFFFFF
    return float(v - 0x100000000 if v & 0x80000000 else v)


def _u32(x):
    if x != x or x in (INF, -INF):
        return 0.0
    return float(int(x) & 0xFFFFFFFF)


def _u8(x):
    if x != x or x in (INF, -INF):
        return 0.0
    return float(int(x) & 0xFF)


_TYPED = {'Float32Array': _f32, 'Float64Array': float, 'Int32Array': _i32, 'Uint32Array': _u32, 'Uint8Array': _u8}


class JSTyped:
    __slots__ = ('a', 'kind', 'conv')

    def __init__(self, kind, a):
        self.kind = kind
        self.conv = _TYPED[kind]
        self.a = a


class Env:
    __slots__ = ('vars', 'parent')

    def __init__(self, parent=None):
        self.vars = {}
        self.parent = parent


class JSFunction:
    __slots__ = ('name', 'params', 'body', 'env', 'arrow', 'is_async', 'expr_body', 'hoisted', 'interp', 'props',
                 'simple')

    def __init__(self, interp, name, params, body, env, arrow, is_async, expr_body, hoisted):
        self.interp, self.name, self.params, self.body, self.env = interp, name, params, body, env
        self.arrow, self.is_async, self.expr_body, self.hoisted = arrow, is_async, expr_body, hoisted
        self.props = None
        self.simple = all(d is None and not r for _, d, r in params)


class Native:
    __slots__ = ('fn', 'name', 'props')

    def __init__(self, fn, name='native'):
        self.fn = fn          # fn(this, args) -> value
        self.name = name
        self.props = None


class Ret:
    __slots__ = ('v',)

    def __init__(self, v):
        self.v = v


BREAK = object()
CONTINUE = object()


# ---------------------------------------------------------------------------------------------- conversions
def truthy(v):
    if v is True:
        return True
    if v is False or v is None or v is UNDEF:
        return False
    t = type(v)
    if t is float:
        return v == v and v != 0.0
    if t is str:
        return len(v) > 0
    return True


def to_num(v):
    t = type(v)
    if t is float:
        return v
    if t is bool:
        return 1.0 if v else 0.0
    if v is None:
        return 0.0
    if v is UNDEF:
        return NAN
    if t is str:
        s = v.strip()
        if s == '':
            return 0.0
        try:
            return float(int(s, 16)) if s[:2] in ('0x', '0X') else float(s)
        except ValueError:
            return NAN
    if t is JSArray:
        if len(v.a) == 0:
            return 0.0
        if len(v.a) == 1:
            return to_num(v.a[0])
    return NAN


def num_to_str(x):
    if x != x:
        return 'NaN'
    if x == INF:
        return 'Infinity'
    if x == -INF:
        return '-Infinity'
    if x == int(x) and abs(x) < 1e21:
        return str(int(x))
    r = repr(x)
    if 'e' in r:
        m, e = r.split('e')
        ex = int(e)
        if m.endswith('.0'):
            m = m[:-2]
        if -7 < ex < 21:                      # JS prints these positionally
            return format(Decimal(r), 'f')
        return m + 'e' + ('+' if ex > 0 else '-') + str(abs(ex))
    return r


def to_str(v):
    t = type(v)
    if t is str:
        return v
    if t is float:
        return num_to_str(v)
    if t is bool:
        return 'true' if v else 'false'
    if v is None:
        return 'null'
    if v is UNDEF:
        return 'undefined'
    if t is JSArray:
        return ','.join('' if (e is None or e is UNDEF) else to_str(e) for e in v.a)
    if t is JSTyped:
        return ','.join(num_to_str(e) for e in v.a)
    if t is JSObject:
        if 'message' in v.props and 'name' in v.props:
            return to_str(v.props['name']) + ': ' + to_str(v.props['message'])
        return '[object Object]'
    if t in (JSFunction, Native):
        return 'function ' + (v.name or '') + '() { [code] }'
    if t is JSPromise:
        return '[object Promise]'
    return str(v)


def prop_key(k):
    """Property key for non-index access: canonical string."""
    return k if type(k) is str else to_str(k)


def loose_eq(a, b):
    ta, tb = type(a), type(b)
    if ta is tb:
        if ta is float or ta is str or ta is bool:
            return a == b
        return a is b
    an = a is None or a is UNDEF
    bn = b is None or b is UNDEF
    if an or bn:
        return an and bn
    if ta in (float, bool, str) and tb in (float, bool, str):
        return to_num(a) == to_num(b)
    if ta in (float, bool, str):
        return loose_eq(a, to_str(b))
    if tb in (float, bool, str):
        return loose_eq(to_str(a), b)
    return a is b


def strict_eq(a, b):
    ta, tb = type(a), type(b)
    if ta is not tb:
        return False
    if ta is float or ta is str or ta is bool:
        return a == b
    return a is b


def js_typeof(v):
    t = type(v)
    if t is float:
        return 'number'
    if t is str:
        return 'string'
    if t is bool:
        return 'boolean'
    if v is UNDEF:
        return 'undefined'
    if t in (JSFunction, Native):
        return 'function'
    return 'object'


def to_int32(v):
    return int(_i32(to_num(v)))


def js_add(a, b):
    if type(a) is float and type(b) is float:
        return a + b
    if type(a) in (JSArray, JSObject, JSTyped, JSFunction, Native):
        a = to_str(a)
    if type(b) in (JSArray, JSObject, JSTyped, JSFunction, Native):
        b = to_str(b)
    if type(a) is str or type(b) is str:
        return to_str(a) + to_str(b)
    return to_num(a) + to_num(b)


def js_div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return NAN
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -INF if neg else INF


def js_mod(a, b):
    if b == 0.0 or a != a or b != b or a in (INF, -INF):
        return NAN
    if b in (INF, -INF):
        return a
    return math.fmod(a, b)


def js_pow(a, b):
    return jsmath.pow(a, b)


def js_less(a, b):          # a < b, undefined (NaN involved) -> None
    if type(a) is float and type(b) is float:
        return a < b
    if type(a) is str and type(b) is str:
        return a < b
    return to_num(a) < to_num(b)


def to_fixed(x, digits):
    """Number.prototype.toFixed: exact decimal value of the double, ties away from zero on the magnitude."""
    if x != x:
        return 'NaN'
    if abs(x) >= 1e21:
        return num_to_str(x)
    q = Decimal(1).scaleb(-digits)
    d = Decimal(abs(x)).quantize(q, rounding=ROUND_HALF_UP)
    s = format(d, 'f')
    return '-' + s if x < 0 and d != 0 else s


def js_parse_int(v, radix=UNDEF):
    s = to_str(v).strip()
    base = 10 if radix is UNDEF else int(to_num(radix))
    sign = 1
    if s[:1] in ('+', '-'):
        sign = -1 if s[0] == '-' else 1
        s = s[1:]
    if (base in (0, 10, 16)) and s[:2] in ('0x', '0X'):
        base, s = 16, s[2:]
    if base == 0:
        base = 10
    digits = '0123456789abcdefghijklmnopqrstuvwxyz'[:base]
    n = 0
    while n < len(s) and s[n].lower() in digits:
        n += 1
    if n == 0:
        return NAN
    return float(sign * int(s[:n], base))


# ---------------------------------------------------------------------------------------------- promises
class JSPromise:
    __slots__ = ('state', 'value', 'reactions', 'interp')

    def __init__(self, interp):
        self.interp = interp
        self.state = 0            # 0 pending, 1 fulfilled, 2 rejected
        self.value = UNDEF
        self.reactions = []

    def resolve(self, v):
        if self.state:
            return
        if type(v) is JSPromise:
            if v is self:
                return self.reject(type_error('promise resolved with itself'))
            # adopt (takes one extra micro-task like the spec's NewPromiseResolveThenableJob)
            self.interp.microtasks.append(lambda: v._subscribe(self.resolve, self.reject))
            return
        self.state, self.value = 1, v
        self._flush()

    def reject(self, v):
        if self.state:
            return
        self.state, self.value = 2, v
        self._flush()

    def _flush(self):
        rs, self.reactions = self.reactions, []
        for on_ok, on_err in rs:
            self._schedule(on_ok, on_err)

    def _schedule(self, on_ok, on_err):
        cb = on_ok if self.state == 1 else on_err
        val = self.value
        self.interp.microtasks.append(lambda: cb(val))

    def _subscribe(self, on_ok, on_err):
        if self.state:
            self._schedule(on_ok, on_err)
        else:
            self.reactions.append((on_ok, on_err))

    def then(self, on_ok, on_err):
        interp = self.interp
        child = JSPromise(interp)

        def ok(v):
            if type(on_ok) in (JSFunction, Native):
                try:
                    child.resolve(interp.call(on_ok, UNDEF, [v]))
                except JSThrow as e:
                    child.reject(e.value)
            else:
                child.resolve(v)

        def err(v):
            if type(on_err) in (JSFunction, Native):
                try:
                    child.resolve(interp.call(on_err, UNDEF, [v]))
                except JSThrow as e:
                    child.reject(e.value)
            else:
                child.reject(v)

        self._subscribe(ok, err)
        return child


# ---------------------------------------------------------------------------------------------- interpreter
class Interp:
    def __init__(self):
        self.microtasks = []
        self.timers = []
        self.console = []                      # (level, text)
        self.globals = Env()
        self._array_proto = self._make_array_proto()
        self._typed_proto = self._make_typed_proto()
        self._promise_proto = self._make_promise_proto()
        self._number_proto = {'toFixed': Native(lambda this, a: to_fixed(this, int(to_num(a[0])) if a else 0), 'toFixed'),
                              'toString': Native(lambda this, a: num_to_str(this), 'toString')}
        self._string_proto = {
            'indexOf': Native(lambda this, a: float(this.find(to_str(a[0]))), 'indexOf'),
            'slice': Native(lambda this, a: this[_slice_idx(a, 0, len(this)):_slice_idx(a, 1, len(this))], 'slice'),
            'toString': Native(lambda this, a: this, 'toString'),
        }
        self._function_proto = {
            'call': Native(lambda this, a: self.call(this, a[0] if a else UNDEF, list(a[1:])), 'call'),
            'apply': Native(lambda this, a: self.call(this, a[0] if a else UNDEF,
                                                      list(a[1].a) if len(a) > 1 and type(a[1]) is JSArray else []), 'apply'),
            'bind': Native(lambda this, a: Native(lambda _t, b, f=this, t=(a[0] if a else UNDEF), pre=list(a[1:]):
                                                  self.call(f, t, pre + list(b)), 'bound'), 'bind'),
        }
        self._install_globals()

    # -- event loop pieces (driven explicitly by the harness)
    def run_microtasks(self):
        q = self.microtasks
        while q:
            q.pop(0)()

    def run_timers(self):
        """Runs every timer queued so far (in delay order, FIFO for equal delays), with micro-tasks after each."""
        while self.timers:
            self.timers.sort(key=lambda t: t[0])
            _, _, fn = self.timers.pop(0)
            self.call(fn, UNDEF, [])
            self.run_microtasks()

    # -- public helpers
    def eval_expression(self, src, env=None):
        node = Parser(src).parse_expression_only()
        return self.cexpr(node, _Scope(None))(env or self.globals)

    def run(self, src, env=None):
        node = Parser(src).parse_program()
        sc = _Scope(None)
        fn = self.cblock_body(node[1], sc, function_level=True)
        return fn(env or self.globals)

    def call(self, fn, this, args):
        t = type(fn)
        if t is JSFunction:
            return self._call_js(fn, this, args)
        if t is Native:
            return fn.fn(this, args)
        raise type_error(to_str(fn) + ' is not a function')

    def _call_js(self, fn, this, args):
        env = Env(fn.env)
        v = env.vars
        if not fn.arrow:
            v['this'] = this          # (`arguments` is not provided: none of the evaluated modules reads it)
        if fn.name and not fn.arrow and fn.name not in v:
            v[fn.name] = fn
        n = len(args)
        if fn.simple:
            i = 0
            for name, _, _ in fn.params:
                v[name] = args[i] if i < n else UNDEF
                i += 1
        else:
            for i, (name, default, rest) in enumerate(fn.params):
                if rest:
                    v[name] = JSArray(list(args[i:]))
                    break
                val = args[i] if i < n else UNDEF
                if val is UNDEF and default is not None:
                    val = default(env)
                v[name] = val
        if fn.is_async:
            p = JSPromise(self)
            try:
                p.resolve(self._run_body(fn, env))
            except JSThrow as e:
                p.reject(e.value)
            return p
        return self._run_body(fn, env)

    def _run_body(self, fn, env):
        if fn.expr_body:
            return fn.body(env)
        for name in fn.hoisted:
            if name not in env.vars:
                env.vars[name] = UNDEF
        c = fn.body(env)
        if type(c) is Ret:
            return c.v
        return UNDEF

    def construct(self, fn, args):
        t = type(fn)
        if t is Native:
            return fn.fn(None, args)          # native constructors ignore `this`
        if t is JSFunction:
            obj = JSObject()
            r = self._call_js(fn, obj, args)
            return r if type(r) in (JSObject, JSArray, JSTyped, JSFunction) else obj
        raise type_error(to_str(fn) + ' is not a constructor')

    # -- property access
    def get(self, obj, key):
        t = type(obj)
        if t is JSArray:
            if type(key) is float:
                i = int(key) if key == key and key not in (INF, -INF) else -1
                if i == key and 0 <= i < len(obj.a):
                    return obj.a[i]
                if i == key and i >= 0:
                    return UNDEF
                key = num_to_str(key)
            if key == 'length':
                return float(len(obj.a))
            if type(key) is str and key.isdigit():
                i = int(key)
                return obj.a[i] if i < len(obj.a) else UNDEF
            k = prop_key(key)
            if obj.props is not None and k in obj.props:
                return obj.props[k]
            return self._array_proto.get(k, UNDEF)
        if t is JSTyped:
            if type(key) is float:
                i = int(key) if key == key and key not in (INF, -INF) else -1
                if i == key and 0 <= i < len(obj.a):
                    return obj.a[i]
                return UNDEF
            if key == 'length':
                return float(len(obj.a))
            if type(key) is str and key.isdigit():
                i = int(key)
                return obj.a[i] if i < len(obj.a) else UNDEF
            return self._typed_proto.get(prop_key(key), UNDEF)
        if t is JSObject:
            k = prop_key(key)
            if obj.getters is not None and k in obj.getters:
                return self.call(obj.getters[k], obj, [])
            return obj.props.get(k, UNDEF)
        if obj is UNDEF or obj is None:
            raise type_error(f"Cannot read properties of {to_str(obj)} (reading '{prop_key(key)}')")
        if t is float:
            return self._number_proto.get(prop_key(key), UNDEF)
        if t is str:
            if type(key) is float:
                i = int(key)
                return obj[i] if 0 <= i < len(obj) and i == key else UNDEF
            if key == 'length':
                return float(len(obj))
            return self._string_proto.get(prop_key(key), UNDEF)
        if t in (JSFunction, Native):
            k = prop_key(key)
            if obj.props is not None and k in obj.props:
                return obj.props[k]
            return self._function_proto.get(k, UNDEF)
        if t is JSPromise:
            return self._promise_proto.get(prop_key(key), UNDEF)
        if t is bool:
            return UNDEF
        raise type_error('unsupported receiver ' + repr(obj))

    def put(self, obj, key, val):
        t = type(obj)
        if t is JSArray:
            if type(key) is float:
                i = int(key) if key == key and key not in (INF, -INF) else -1
                if i == key and i >= 0:
                    a = obj.a
                    n = len(a)
                    if i < n:
                        a[i] = val
                    elif i == n:
                        a.append(val)
                    else:
                        a.extend([UNDEF] * (i - n))
                        a.append(val)
                    return
            elif type(key) is str and key.isdigit():
                return self.put(obj, float(int(key)), val)
            elif key == 'length':
                n = int(to_num(val))
                if n < len(obj.a):
                    del obj.a[n:]
                else:
                    obj.a.extend([UNDEF] * (n - len(obj.a)))
                return
            if obj.props is None:
                obj.props = {}
            obj.props[prop_key(key)] = val
            return
        if t is JSTyped:
            if type(key) is float:
                i = int(key)
                if i == key and 0 <= i < len(obj.a):
                    obj.a[i] = obj.conv(to_num(val))
                return
            if type(key) is str and key.isdigit():
                return self.put(obj, float(int(key)), val)
            return
        if t is JSObject:
            obj.props[prop_key(key)] = val
            return
        if t in (JSFunction, Native):
            if obj.props is None:
                obj.props = {}
            obj.props[prop_key(key)] = val
            return
        if obj is UNDEF or obj is None:
            raise type_error(f"Cannot set properties of {to_str(obj)} (setting '{prop_key(key)}')")
        # primitives: silently ignored in sloppy mode, TypeError in strict; the modules never do it

    # -- built-ins
    def _make_array_proto(self):
        call = self.call

        def push(this, a):
            this.a.extend(a)
            return float(len(this.a))

        def pop(this, a):
            return this.a.pop() if this.a else UNDEF

        def shift(this, a):
            return this.a.pop(0) if this.a else UNDEF

        def slice_(this, a):
            n = len(this.a)
            return JSArray(this.a[_slice_idx(a, 0, n):_slice_idx(a, 1, n)])

        def splice(this, a):
            n = len(this.a)
            start = _slice_idx(a, 0, n)
            cnt = n - start if len(a) < 2 else max(0, min(int(to_num(a[1])), n - start))
            removed = this.a[start:start + cnt]
            this.a[start:start + cnt] = list(a[2:])
            return JSArray(removed)

        def concat(this, a):
            out = list(this.a)
            for x in a:
                if type(x) is JSArray:
                    out.extend(x.a)
                else:
                    out.append(x)          # typed arrays are not spreadable: appended as one element
            return JSArray(out)

        def reduce(this, a):
            fn = a[0]
            items = this.a
            if len(a) > 1:
                acc, i = a[1], 0
            else:
                if not items:
                    raise type_error('Reduce of empty array with no initial value')
                acc, i = items[0], 1
            for k in range(i, len(items)):
                acc = call(fn, UNDEF, [acc, items[k], float(k), this])
            return acc

        def fill(this, a):
            v = a[0] if a else UNDEF
            n = len(this.a)
            s = _slice_idx(a, 1, n) if len(a) > 1 else 0
            e = _slice_idx(a, 2, n) if len(a) > 2 else n
            for i in range(s, e):
                this.a[i] = v
            return this

        def index_of(this, a):
            for i, v in enumerate(this.a):
                if strict_eq(v, a[0]):
                    return float(i)
            return -1.0

        def map_(this, a):
            return JSArray([call(a[0], UNDEF, [v, float(i), this]) for i, v in enumerate(list(this.a))])

        def for_each(this, a):
            for i, v in enumerate(list(this.a)):
                call(a[0], UNDEF, [v, float(i), this])
            return UNDEF

        def filter_(this, a):
            return JSArray([v for i, v in enumerate(list(this.a)) if truthy(call(a[0], UNDEF, [v, float(i), this]))])

        def find_index(this, a):
            for i, v in enumerate(list(this.a)):
                if truthy(call(a[0], UNDEF, [v, float(i), this])):
                    return float(i)
            return -1.0

        def join(this, a):
            sep = ',' if not a or a[0] is UNDEF else to_str(a[0])
            return sep.join('' if (e is None or e is UNDEF) else to_str(e) for e in this.a)

        def sort(this, a):
            import functools
            if a and type(a[0]) in (JSFunction, Native):
                def cmp(x, y):
                    r = to_num(call(a[0], UNDEF, [x, y]))
                    return -1 if r < 0 else (1 if r > 0 else 0)
            else:
                def cmp(x, y):
                    sx, sy = to_str(x), to_str(y)
                    return -1 if sx < sy else (1 if sx > sy else 0)
            this.a.sort(key=functools.cmp_to_key(cmp))
            return this

        return {k: Native(f, k) for k, f in dict(
            push=push, pop=pop, shift=shift, slice=slice_, splice=splice, concat=concat, reduce=reduce, fill=fill,
            indexOf=index_of, findIndex=find_index, map=map_, forEach=for_each, filter=filter_, join=join, sort=sort,
            toString=lambda this, a: to_str(this)).items()}

    def _make_typed_proto(self):
        def fill(this, a):
            v = this.conv(to_num(a[0] if a else UNDEF))
            for i in range(len(this.a)):
                this.a[i] = v
            return this

        def slice_(this, a):
            n = len(this.a)
            return JSTyped(this.kind, this.a[_slice_idx(a, 0, n):_slice_idx(a, 1, n)])

        # subarray: a view in JS; the callers here only read it, so a copy is equivalent
        return {'fill': Native(fill, 'fill'), 'slice': Native(slice_, 'slice'), 'subarray': Native(slice_, 'subarray')}

    def _make_promise_proto(self):
        return {
            'then': Native(lambda this, a: this.then(a[0] if a else UNDEF, a[1] if len(a) > 1 else UNDEF), 'then'),
            'catch': Native(lambda this, a: this.then(UNDEF, a[0] if a else UNDEF), 'catch'),
        }

    def _install_globals(self):
        g = self.globals.vars
        nat = Native

        def mfun(f):
            return nat(lambda this, a: f(*[to_num(x) for x in a]))

        def m_sqrt(x=NAN):
            return math.sqrt(x) if x >= 0 else (NAN if x == x else NAN)

        def m_round(x=NAN):
            if x != x or x in (INF, -INF):
                return x
            return float(math.floor(x + 0.5))

        def m_max(*xs):
            r = -INF
            for x in xs:
                if x != x:
                    return NAN
                if x > r:
                    r = x
            return r

        def m_min(*xs):
            r = INF
            for x in xs:
                if x != x:
                    return NAN
                if x < r:
                    r = x
            return r

        def m_log10(x=NAN):
            if x != x or x < 0:
                return NAN
            if x == 0:
                return -INF
            return jsmath.log10(x)

        def m_log(x=NAN):
            if x != x or x < 0:
                return NAN
            if x == 0:
                return -INF
            return jsmath.log(x)

        g['Math'] = JSObject({
            'abs': mfun(lambda x=NAN: abs(x)), 'sqrt': mfun(m_sqrt), 'pow': mfun(lambda x=NAN, y=NAN: js_pow(x, y)),
            'log10': mfun(m_log10), 'log': mfun(m_log), 'floor': mfun(lambda x=NAN: float(math.floor(x)) if math.isfinite(x) else x),
            'ceil': mfun(lambda x=NAN: float(math.ceil(x)) if math.isfinite(x) else x), 'round': mfun(m_round),
            'max': mfun(m_max), 'min': mfun(m_min), 'PI': math.pi, 'E': math.e,
        })
        g['parseInt'] = nat(lambda this, a: js_parse_int(a[0] if a else UNDEF, a[1] if len(a) > 1 else UNDEF), 'parseInt')
        g['parseFloat'] = nat(lambda this, a: to_num(a[0] if a else UNDEF), 'parseFloat')
        g['isNaN'] = nat(lambda this, a: to_num(a[0] if a else UNDEF) != to_num(a[0] if a else UNDEF), 'isNaN')
        g['Number'] = nat(lambda this, a: to_num(a[0]) if a else 0.0, 'Number')
        g['String'] = nat(lambda this, a: to_str(a[0]) if a else '', 'String')
        g['NaN'] = NAN
        g['Infinity'] = INF
        g['undefined'] = UNDEF

        def array_ctor(this, a):
            if len(a) == 1 and type(a[0]) is float:
                return JSArray([UNDEF] * int(a[0]))
            return JSArray(list(a))
        arr = nat(array_ctor, 'Array')
        arr.props = {'isArray': nat(lambda this, a: type(a[0]) is JSArray if a else False, 'isArray'),
                     # Array.from(arrayLike): a plain array of the elements (the Node shim copies typed-array rows with it)
                     'from': nat(lambda this, a: JSArray(list(a[0].a)) if a and type(a[0]) in (JSArray, JSTyped) else JSArray([]), 'from')}
        g['Array'] = arr

        def typed_ctor(kind):
            conv = _TYPED[kind]

            def ctor(this, a):
                if not a:
                    return JSTyped(kind, [])
                x = a[0]
                if type(x) is float:
                    return JSTyped(kind, [0.0] * int(x))
                if type(x) in (JSArray, JSTyped):
                    return JSTyped(kind, [conv(to_num(v)) for v in x.a])
                raise type_error('minijs: unsupported typed-array constructor argument')
            return nat(ctor, kind)
        for kind in _TYPED:
            g[kind] = typed_ctor(kind)

        def promise_ctor(this, a):
            p = JSPromise(self)
            try:
                self.call(a[0], UNDEF, [nat(lambda t, b: p.resolve(b[0] if b else UNDEF), 'resolve'),
                                        nat(lambda t, b: p.reject(b[0] if b else UNDEF), 'reject')])
            except JSThrow as e:          # a throw inside the executor rejects the promise
                p.reject(e.value)
            return p
        pr = nat(promise_ctor, 'Promise')

        def p_resolve(this, a):
            p = JSPromise(self)
            p.resolve(a[0] if a else UNDEF)
            return p
        def p_all(this, a):
            # Promise.all over an array: resolves (one micro-task after the last element) with the values in order, rejects
            # with the first rejection (src/prediction.js:65 waits for one nn_prediction per model this way)
            items = list(a[0].a) if a and isinstance(a[0], JSArray) else []
            out = JSPromise(self)
            vals = [UNDEF] * len(items)
            left = [len(items)]
            if not items:
                out.resolve(JSArray([]))
                return out

            def settle(i):
                def ok(v):
                    vals[i] = v
                    left[0] -= 1
                    if left[0] == 0:
                        out.resolve(JSArray(list(vals)))
                return ok
            for i, it_ in enumerate(items):
                q = it_ if type(it_) is JSPromise else p_resolve(UNDEF, [it_])
                q._subscribe(settle(i), out.reject)
            return out
        pr.props = {'resolve': nat(p_resolve, 'resolve'), 'all': nat(p_all, 'all')}
        g['Promise'] = pr

        def logger(level):
            def f(this, a):
                self.console.append((level, ' '.join(to_str(x) for x in a)))
                return UNDEF
            return nat(f, level)
        g['console'] = JSObject({k: logger(k) for k in ('log', 'error', 'warn', 'info')})

        def set_timeout(this, a):
            self.timers.append((to_num(a[1]) if len(a) > 1 else 0.0, len(self.timers), a[0]))
            return float(len(self.timers))
        st = nat(set_timeout, 'setTimeout')
        g['setTimeout'] = st
        g['window'] = JSObject({'setTimeout': st})
        g['self'] = g['window']

        def define_property(this, a):
            obj, name, desc = a[0], to_str(a[1]), a[2]
            getter = self.get(desc, 'get')
            if getter is not UNDEF:
                if obj.getters is None:
                    obj.getters = {}
                obj.getters[name] = getter
            else:
                self.put(obj, name, self.get(desc, 'value'))
            return obj

        def keys(this, a):
            o = a[0]
            if type(o) is JSObject:
                return JSArray(list(o.props.keys()))
            return JSArray([str(i) for i in range(len(o.a))])
        g['Object'] = JSObject({'defineProperty': nat(define_property, 'defineProperty'), 'keys': nat(keys, 'keys')})

    # ------------------------------------------------------------------------------------------ compiler
    def cblock_body(self, stmts, scope, function_level=False):
        """Compiles a statement list; function declarations are hoisted to the start of the list."""
        fdecls = [s for s in stmts if s[0] == 'fdecl']
        others = [s for s in stmts if s[0] != 'fdecl']
        hoist = [(s[1], self.cexpr(s[2], scope)) for s in fdecls]
        body = [self.cstmt(s, scope) for s in others]

        def run(env):
            v = env.vars
            for name, mk in hoist:
                v[name] = mk(env)
            for st in body:
                c = st(env)
                if c is not None:
                    return c
            return None
        return run

    @staticmethod
    def _declares(stmts):
        for s in stmts:
            if s[0] == 'fdecl' or (s[0] == 'var' and s[1] in ('let', 'const')):
                return True
        return False

    def cstmt(self, node, scope):
        k = node[0]
        if k == 'expr':
            e = self.cexpr(node[1], scope)

            def s_expr(env):
                e(env)
            return s_expr
        if k == 'var':
            kind = node[1]
            decls = [(name, self.cexpr(init, scope) if init is not None else None) for name, init in node[2]]
            if kind == 'var':
                for name, _ in decls:
                    scope.declare_var(name)

                def s_var(env):
                    for name, init in decls:
                        if init is not None:
                            self._assign_name(env, name, init(env))
                return s_var

            def s_let(env):
                v = env.vars
                for name, init in decls:
                    v[name] = init(env) if init is not None else UNDEF
            return s_let
        if k == 'block':
            inner = self.cblock_body(node[1], scope)
            if self._declares(node[1]):
                return lambda env: inner(Env(env))
            return inner
        if k == 'if':
            test, cons = self.cexpr(node[1], scope), self.cstmt(node[2], scope)
            alt = self.cstmt(node[3], scope) if node[3] is not None else None
            if alt is None:
                def s_if(env):
                    if truthy(test(env)):
                        return cons(env)
                return s_if

            def s_ifelse(env):
                if truthy(test(env)):
                    return cons(env)
                return alt(env)
            return s_ifelse
        if k == 'for':
            init = self.cstmt(node[1], scope) if node[1] is not None else None
            test = self.cexpr(node[2], scope) if node[2] is not None else None
            update = self.cexpr(node[3], scope) if node[3] is not None else None
            body = self.cstmt(node[4], scope)
            new_env = node[1] is not None and node[1][0] == 'var' and node[1][1] != 'var'

            def s_for(env):
                if new_env:
                    env = Env(env)
                if init is not None:
                    init(env)
                while test is None or truthy(test(env)):
                    c = body(env)
                    if c is not None:
                        if c is BREAK:
                            break
                        if c is not CONTINUE:
                            return c
                    if update is not None:
                        update(env)
            return s_for
        if k == 'while':
            test, body = self.cexpr(node[1], scope), self.cstmt(node[2], scope)

            def s_while(env):
                while truthy(test(env)):
                    c = body(env)
                    if c is not None:
                        if c is BREAK:
                            break
                        if c is not CONTINUE:
                            return c
            return s_while
        if k == 'dowhile':
            body, test = self.cstmt(node[1], scope), self.cexpr(node[2], scope)

            def s_dowhile(env):
                while True:
                    c = body(env)
                    if c is not None:
                        if c is BREAK:
                            break
                        if c is not CONTINUE:
                            return c
                    if not truthy(test(env)):
                        break
            return s_dowhile
        if k in ('forin', 'forof'):
            kind, name, obj, body = node[1], node[2], self.cexpr(node[3], scope), self.cstmt(node[4], scope)
            if kind == 'var':
                scope.declare_var(name)
            is_in = k == 'forin'

            def s_forin(env):
                o = obj(env)
                t = type(o)
                if is_in:
                    if t in (JSArray, JSTyped):
                        keys = [str(i) for i, v in enumerate(o.a)]
                        if t is JSArray and o.props:
                            keys += list(o.props.keys())
                    elif t is JSObject:
                        keys = list(o.props.keys())
                    elif o is None or o is UNDEF:
                        keys = []
                    else:
                        raise type_error('minijs: for-in over ' + js_typeof(o))
                else:
                    if t in (JSArray, JSTyped):
                        keys = list(o.a)
                    else:
                        raise type_error('minijs: for-of over ' + js_typeof(o))
                for key in keys:
                    if kind in ('let', 'const'):
                        e2 = Env(env)
                        e2.vars[name] = key
                    else:
                        e2 = env
                        self._assign_name(env, name, key)
                    c = body(e2)
                    if c is not None:
                        if c is BREAK:
                            break
                        if c is not CONTINUE:
                            return c
            return s_forin
        if k == 'return':
            if node[1] is None:
                r = Ret(UNDEF)
                return lambda env: r
            arg = self.cexpr(node[1], scope)
            return lambda env: Ret(arg(env))
        if k == 'break':
            return lambda env: BREAK
        if k == 'continue':
            return lambda env: CONTINUE
        if k == 'empty':
            return lambda env: None
        if k == 'fdecl':            # only reached when a declaration is the sole body of if/for (not in these modules)
            mk = self.cexpr(node[2], scope)
            name = node[1]

            def s_fdecl(env):
                env.vars[name] = mk(env)
            return s_fdecl
        if k == 'throw':
            arg = self.cexpr(node[1], scope)

            def s_throw(env):
                raise JSThrow(arg(env))
            return s_throw
        if k == 'try':
            blk = self.cstmt(node[1], scope)
            param = node[2]
            handler = self.cstmt(node[3], scope) if node[3] is not None else None
            final = self.cstmt(node[4], scope) if node[4] is not None else None

            def s_try(env):
                try:
                    try:
                        return blk(env)
                    except JSThrow as e:
                        if handler is None:
                            raise
                        e2 = Env(env)
                        if param:
                            e2.vars[param] = e.value
                        return handler(e2)
                finally:
                    if final is not None:
                        c = final(env)
                        if c is not None:
                            return c
            return s_try
        if k == 'switch':
            disc = self.cexpr(node[1], scope)
            cases = [(self.cexpr(t, scope) if t is not None else None, [self.cstmt(s, scope) for s in body])
                     for t, body in node[2]]

            def s_switch(env):
                d = disc(env)
                env = Env(env)
                start = None
                for i, (t, _) in enumerate(cases):
                    if t is not None and strict_eq(d, t(env)):
                        start = i
                        break
                if start is None:
                    for i, (t, _) in enumerate(cases):
                        if t is None:
                            start = i
                            break
                if start is None:
                    return None
                for _, body in cases[start:]:
                    for st in body:
                        c = st(env)
                        if c is not None:
                            if c is BREAK:
                                return None
                            return c
            return s_switch
        raise SyntaxError('minijs: statement kind ' + k)

    def _assign_name(self, env, name, val):
        e = env
        while e is not None:
            if name in e.vars:
                e.vars[name] = val
                return
            e = e.parent
        self.globals.vars[name] = val          # sloppy-mode implicit global

    def cexpr(self, node, scope):
        k = node[0]
        get, put, call = self.get, self.put, self.call
        if k == 'num' or k == 'str' or k == 'const':
            v = node[1]
            return lambda env: v
        if k == 'id':
            name = node[1]

            def e_id(env):
                e = env
                while e is not None:
                    v = e.vars
                    if name in v:
                        return v[name]
                    e = e.parent
                raise JSThrow(JSObject({'name': 'ReferenceError', 'message': name + ' is not defined'}))
            return e_id
        if k == 'this':
            def e_this(env):
                e = env
                while e is not None:
                    if 'this' in e.vars:
                        return e.vars['this']
                    e = e.parent
                return UNDEF
            return e_this
        if k == 'member':
            obj, name = self.cexpr(node[1], scope), node[2]
            return lambda env: get(obj(env), name)
        if k == 'index':
            obj, idx = self.cexpr(node[1], scope), self.cexpr(node[2], scope)

            def e_index(env):
                o = obj(env)
                i = idx(env)
                # fast paths
                if type(i) is float:
                    t = type(o)
                    if t is JSArray or t is JSTyped:
                        a = o.a
                        try:
                            ii = int(i)
                        except (ValueError, OverflowError):
                            return get(o, i)
                        if ii == i and 0 <= ii < len(a):
                            return a[ii]
                return get(o, i)
            return e_index
        if k == 'seq':
            items = [self.cexpr(x, scope) for x in node[1]]
            last = items[-1]
            first = items[:-1]

            def e_seq(env):
                for it in first:
                    it(env)
                return last(env)
            return e_seq
        if k == 'cond':
            t, a, b = self.cexpr(node[1], scope), self.cexpr(node[2], scope), self.cexpr(node[3], scope)
            return lambda env: a(env) if truthy(t(env)) else b(env)
        if k == 'logical':
            l, r = self.cexpr(node[2], scope), self.cexpr(node[3], scope)
            if node[1] == '&&':
                def e_and(env):
                    v = l(env)
                    return r(env) if truthy(v) else v
                return e_and

            def e_or(env):
                v = l(env)
                return v if truthy(v) else r(env)
            return e_or
        if k == 'bin':
            return self._cbin(node[1], self.cexpr(node[2], scope), self.cexpr(node[3], scope))
        if k == 'unary':
            op, a = node[1], None
            if op == 'typeof' and node[2][0] == 'id':
                name = node[2][1]

                def e_typeof_id(env):
                    e = env
                    while e is not None:
                        if name in e.vars:
                            return js_typeof(e.vars[name])
                        e = e.parent
                    return 'undefined'
                return e_typeof_id
            if op == 'delete':
                raise SyntaxError('minijs: delete is not supported')
            a = self.cexpr(node[2], scope)
            if op == '!':
                return lambda env: not truthy(a(env))
            if op == '-':
                return lambda env: -to_num(a(env))
            if op == '+':
                return lambda env: to_num(a(env))
            if op == '~':
                return lambda env: float(~to_int32(a(env)))
            if op == 'typeof':
                return lambda env: js_typeof(a(env))
            if op == 'void':
                def e_void(env):
                    a(env)
                    return UNDEF
                return e_void
        if k == 'assign':
            return self._cassign(node, scope)
        if k == 'update':
            delta = 1.0 if node[1] == '++' else -1.0
            prefix, target = node[2], node[3]
            ref_get, ref_set = self._cref(target, scope)

            def e_update(env):
                ctx = ref_get(env)
                old = to_num(ctx[0])
                new = old + delta
                ref_set(env, ctx, new)
                return new if prefix else old
            return e_update
        if k == 'call':
            callee, args = node[1], [self.cexpr(a, scope) for a in node[2]]
            if callee[0] in ('member', 'index'):
                obj = self.cexpr(callee[1], scope)
                if callee[0] == 'member':
                    name = callee[2]
                    key = lambda env: name       # noqa: E731
                else:
                    key = self.cexpr(callee[2], scope)

                def e_mcall(env):
                    o = obj(env)
                    f = get(o, key(env))
                    argv = [a(env) for a in args]
                    t = type(f)
                    if t is Native:
                        return f.fn(o, argv)
                    if t is JSFunction:
                        return self._call_js(f, o, argv)
                    raise type_error(prop_key(key(env)) + ' is not a function')
                return e_mcall
            fn = self.cexpr(callee, scope)

            def e_call(env):
                f = fn(env)
                argv = [a(env) for a in args]
                t = type(f)
                if t is JSFunction:
                    return self._call_js(f, UNDEF, argv)
                if t is Native:
                    return f.fn(UNDEF, argv)
                raise type_error(to_str(f) + ' is not a function')
            return e_call
        if k == 'new':
            fn, args = self.cexpr(node[1], scope), [self.cexpr(a, scope) for a in node[2]]
            return lambda env: self.construct(fn(env), [a(env) for a in args])
        if k == 'arr':
            items = [self.cexpr(x, scope) if x is not None else None for x in node[1]]
            return lambda env: JSArray([it(env) if it is not None else UNDEF for it in items])
        if k == 'obj':
            props = [(prop_key(key), self.cexpr(v, scope)) for key, v in node[1]]
            return lambda env: JSObject({key: v(env) for key, v in props})
        if k == 'fn':
            _, name, params, body, arrow, is_async, expr_body = node
            inner = _Scope(scope)
            cparams = [(n, self.cexpr(d, inner) if d is not None else None, r) for n, d, r in params]
            if expr_body:
                cbody = self.cexpr(body, inner)
            else:
                cbody = self.cblock_body(body[1], inner, function_level=True)
            hoisted = tuple(inner.vars)
            return lambda env: JSFunction(self, name, cparams, cbody, env, arrow, is_async, expr_body, hoisted)
        raise SyntaxError('minijs: expression kind ' + k)

    def _cbin(self, op, l, r):
        if op == '+':
            def e_add(env):
                a, b = l(env), r(env)
                if type(a) is float and type(b) is float:
                    return a + b
                return js_add(a, b)
            return e_add
        if op == '-':
            def e_sub(env):
                a, b = l(env), r(env)
                if type(a) is float and type(b) is float:
                    return a - b
                return to_num(a) - to_num(b)
            return e_sub
        if op == '*':
            def e_mul(env):
                a, b = l(env), r(env)
                if type(a) is float and type(b) is float:
                    return a * b
                return to_num(a) * to_num(b)
            return e_mul
        if op == '/':
            return lambda env: js_div(to_num(l(env)), to_num(r(env)))
        if op == '%':
            return lambda env: js_mod(to_num(l(env)), to_num(r(env)))
        if op == '**':
            return lambda env: js_pow(to_num(l(env)), to_num(r(env)))
        if op == '<':
            def e_lt(env):
                a, b = l(env), r(env)
                if type(a) is float and type(b) is float:
                    return a < b
                return js_less(a, b)
            return e_lt
        if op == '>':
            def e_gt(env):
                a, b = l(env), r(env)
                if type(a) is float and type(b) is float:
                    return a > b
                return js_less(b, a)
            return e_gt
        if op == '<=':
            def e_le(env):
                a, b = l(env), r(env)
                if type(a) is float and type(b) is float:
                    return a <= b
                if type(a) is str and type(b) is str:
                    return a <= b
                return to_num(a) <= to_num(b)
            return e_le
        if op == '>=':
            def e_ge(env):
                a, b = l(env), r(env)
                if type(a) is float and type(b) is float:
                    return a >= b
                if type(a) is str and type(b) is str:
                    return a >= b
                return to_num(a) >= to_num(b)
            return e_ge
        if op == '==':
            return lambda env: loose_eq(l(env), r(env))
        if op == '!=':
            return lambda env: not loose_eq(l(env), r(env))
        if op == '===':
            return lambda env: strict_eq(l(env), r(env))
        if op == '!==':
            return lambda env: not strict_eq(l(env), r(env))
        if op == '&':
            return lambda env: float(to_int32(l(env)) & to_int32(r(env)))
        if op == '|':
            return lambda env: float(to_int32(l(env)) | to_int32(r(env)))
        if op == '^':
            return lambda env: float(to_int32(l(env)) ^ to_int32(r(env)))
        if op == '<<':
            return lambda env: _i32(float(to_int32(l(env)) << (to_int32(r(env)) & 31)))
        if op == '>>':
            return lambda env: float(to_int32(l(env)) >> (to_int32(r(env)) & 31))
        if op == '>>>':
            return lambda env: float((to_int32(l(env)) & 0xFFFFFFFF) >> (to_int32(r(env)) & 31))
        if op == 'instanceof':
            def e_instanceof(env):
                a, b = l(env), r(env)
                n = b.name if type(b) in (Native, JSFunction) else None
                if n == 'Array':
                    return type(a) is JSArray
                if n in _TYPED:
                    return type(a) is JSTyped and a.kind == n
                if n == 'Promise':
                    return type(a) is JSPromise
                return False
            return e_instanceof
        if op == 'in':
            def e_in(env):
                a, b = l(env), r(env)
                if type(b) is JSObject:
                    return prop_key(a) in b.props
                if type(b) in (JSArray, JSTyped):
                    return 0 <= to_num(a) < len(b.a)
                raise type_error("minijs: 'in' on " + js_typeof(b))
            return e_in
        raise SyntaxError('minijs: operator ' + op)

    def _cref(self, target, scope):
        """(getter(env) -> ctx tuple whose [0] is the current value, setter(env, ctx, value))."""
        get, put = self.get, self.put
        if target[0] == 'id':
            name = target[1]

            def rget(env):
                e = env
                while e is not None:
                    if name in e.vars:
                        return (e.vars[name], e)
                    e = e.parent
                raise JSThrow(JSObject({'name': 'ReferenceError', 'message': name + ' is not defined'}))

            def rset(env, ctx, val):
                ctx[1].vars[name] = val
            return rget, rset
        if target[0] == 'member':
            obj, name = self.cexpr(target[1], scope), target[2]

            def rget(env):
                o = obj(env)
                return (get(o, name), o)

            def rset(env, ctx, val):
                put(ctx[1], name, val)
            return rget, rset
        if target[0] == 'index':
            obj, idx = self.cexpr(target[1], scope), self.cexpr(target[2], scope)

            def rget(env):
                o = obj(env)
                i = idx(env)
                return (get(o, i), o, i)

            def rset(env, ctx, val):
                put(ctx[1], ctx[2], val)
            return rget, rset
        raise SyntaxError('minijs: invalid update/assignment target')

    def _cassign(self, node, scope):
        _, op, target, value = node
        val = self.cexpr(value, scope)
        get, put = self.get, self.put
        if op == '=':
            if target[0] == 'id':
                name = target[1]

                def e_assign_id(env):
                    v = val(env)
                    e = env
                    while e is not None:
                        if name in e.vars:
                            e.vars[name] = v
                            return v
                        e = e.parent
                    self.globals.vars[name] = v
                    return v
                return e_assign_id
            if target[0] == 'member':
                obj, name = self.cexpr(target[1], scope), target[2]

                def e_assign_member(env):
                    o = obj(env)
                    v = val(env)
                    put(o, name, v)
                    return v
                return e_assign_member
            obj, idx = self.cexpr(target[1], scope), self.cexpr(target[2], scope)

            def e_assign_index(env):
                o = obj(env)
                i = idx(env)
                v = val(env)
                put(o, i, v)
                return v
            return e_assign_index
        binop = self._cbin(op[:-1], lambda pair: pair[0], lambda pair: pair[1])   # the "env" is just the operand pair
        ref_get, ref_set = self._cref(target, scope)

        def e_compound(env):
            ctx = ref_get(env)
            new = binop((ctx[0], val(env)))
            ref_set(env, ctx, new)
            # a typed-array store rounds; the expression's value is the unrounded number (as in JS)
            return new
        return e_compound


class _Scope:
    """Compile-time function scope: collects `var` names so they exist (undefined) from function entry."""
    __slots__ = ('vars', 'parent')

    def __init__(self, parent):
        self.vars = []
        self.parent = parent

    def declare_var(self, name):
        if name not in self.vars:
            self.vars.append(name)


def _slice_idx(args, k, n):
    if len(args) <= k or args[k] is UNDEF:
        return 0 if k == 0 else n
    x = to_num(args[k])
    if x != x:
        return 0
    i = int(max(min(x, 1e15), -1e15))
    if i < 0:
        i = max(0, n + i)
    return min(i, n)
