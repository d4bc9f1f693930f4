"""TEST INFRASTRUCTURE ONLY.  Tokenizer + Pratt parser for the ECMAScript subset the reference's minified
formantanalyzer modules use (ES2017 minus classes, generators, destructuring, template and regex literals).
Produces plain tuples; oracle/minijs/interp.py compiles them to closures.  Unsupported syntax raises SyntaxError
-- nothing is silently skipped."""
from __future__ import annotations

import re

_TOKEN_RE = re.compile(r'''
  (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
 |(?P<num>0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)
 |(?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
 |(?P<str>"(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*')
 |(?P<punc>>>>=|\.\.\.|===|!==|\*\*=|<<=|>>=|>>>|=>|==|!=|<=|>=|&&|\|\||\+\+|--|\+=|-=|\*=|/=|%=|&=|\|=|\^=|\*\*|<<|>>|[{}()\[\];,<>+\-*/%&|^!~?:=.])
''', re.X | re.S)

_ESC = {'n': '\n', 't': '\t', 'r': '\r', 'b': '\b', 'f': '\f', 'v': '\v', '0': '\0'}
KEYWORDS = {'var', 'let', 'const', 'function', 'return', 'if', 'else', 'for', 'while', 'do', 'break', 'continue',
            'new', 'this', 'typeof', 'void', 'delete', 'in', 'of', 'instanceof', 'try', 'catch', 'finally', 'throw',
            'true', 'false', 'null', 'async', 'switch', 'case', 'default'}


def _unescape(s: str) -> str:
    out, i = [], 0
    while i < len(s):
        c = s[i]
        if c != '\\':
            out.append(c)
            i += 1
            continue
        c = s[i + 1]
        if c == 'u':
            out.append(chr(int(s[i + 2:i + 6], 16)))
            i += 6
        elif c == 'x':
            out.append(chr(int(s[i + 2:i + 4], 16)))
            i += 4
        else:
            out.append(_ESC.get(c, c))
            i += 2
    return ''.join(out)


def tokenize(src: str):
    """-> list of (kind, value, offset); kind in num / id / str / punc / eof."""
    toks, pos, n = [], 0, len(src)
    while pos < n:
        m = _TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"minijs: cannot tokenize at offset {pos}: {src[pos:pos + 40]!r}")
        kind = m.lastgroup
        if kind != 'ws':
            text = m.group()
            if kind == 'num':
                val = float(int(text, 16)) if text[:2] in ('0x', '0X') else float(text)
            elif kind == 'str':
                val = _unescape(text[1:-1])
            else:
                val = text
            toks.append((kind, val, pos))
        pos = m.end()
    toks.append(('eof', None, n))
    return toks


def matching_end(src: str, start: int) -> int:
    """Offset just past the bracket that closes the bracket at src[start] (string-aware)."""
    depth = 0
    for kind, val, off in tokenize_from(src, start):
        if kind == 'punc':
            if val in '([{':
                depth += 1
            elif val in ')]}':
                depth -= 1
                if depth == 0:
                    return off + 1
    raise SyntaxError("minijs: unbalanced brackets")


def tokenize_from(src: str, pos: int):
    n = len(src)
    while pos < n:
        m = _TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"minijs: cannot tokenize at offset {pos}: {src[pos:pos + 40]!r}")
        if m.lastgroup != 'ws':
            yield m.lastgroup, m.group(), pos
        pos = m.end()


# binary operator precedences (higher binds tighter); ** is right-associative
_BINPREC = {'||': 4, '&&': 5, '|': 6, '^': 7, '&': 8, '==': 9, '!=': 9, '===': 9, '!==': 9,
            '<': 10, '>': 10, '<=': 10, '>=': 10, 'instanceof': 10, 'in': 10,
            '<<': 11, '>>': 11, '>>>': 11, '+': 12, '-': 12, '*': 13, '/': 13, '%': 13, '**': 14}
_ASSIGN = {'=', '+=', '-=', '*=', '/=', '%=', '**=', '<<=', '>>=', '>>>=', '&=', '|=', '^='}


class Parser:
    def __init__(self, src: str):
        self.src = src
        self.t = tokenize(src)
        self.i = 0
        self.no_in = False

    # -- token helpers
    def peek(self, k=0):
        return self.t[self.i + k]

    def at(self, val, k=0):
        tk = self.t[self.i + k]
        return tk[0] in ('punc', 'id') and tk[1] == val

    def eat(self, val):
        if self.at(val):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.eat(val):
            tk = self.peek()
            raise SyntaxError(f"minijs: expected {val!r} at offset {tk[2]}, got {tk[1]!r}: {self.src[tk[2]:tk[2] + 40]!r}")

    def ident(self):
        tk = self.peek()
        if tk[0] != 'id':
            raise SyntaxError(f"minijs: identifier expected at offset {tk[2]}: {self.src[tk[2]:tk[2] + 40]!r}")
        self.i += 1
        return tk[1]

    # -- program / statements
    def parse_program(self):
        body = []
        while self.peek()[0] != 'eof':
            body.append(self.statement())
        return ('block', body)

    def parse_expression_only(self):
        e = self.expression()
        if self.peek()[0] != 'eof':
            tk = self.peek()
            raise SyntaxError(f"minijs: trailing input at offset {tk[2]}")
        return e

    def block(self):
        self.expect('{')
        body = []
        while not self.at('}'):
            body.append(self.statement())
        self.expect('}')
        return ('block', body)

    def semi(self):
        if self.eat(';'):
            return
        if self.at('}') or self.peek()[0] == 'eof':
            return
        # minified code always has explicit separators; anything else is a parse error we want to see
        tk = self.peek()
        raise SyntaxError(f"minijs: ';' expected at offset {tk[2]}: {self.src[tk[2]:tk[2] + 40]!r}")

    def statement(self):
        tk = self.peek()
        if tk[0] == 'punc':
            if tk[1] == '{':
                return self.block()
            if tk[1] == ';':
                self.i += 1
                return ('empty',)
        if tk[0] == 'id':
            v = tk[1]
            if v in ('var', 'let', 'const'):
                d = self.var_decl()
                self.semi()
                return d
            if v == 'function' or (v == 'async' and self.at('function', 1)):
                fn = self.function(is_decl=True)
                return ('fdecl', fn[1], fn)
            if v == 'return':
                self.i += 1
                arg = None
                if not (self.at(';') or self.at('}') or self.peek()[0] == 'eof'):
                    arg = self.expression()
                self.semi()
                return ('return', arg)
            if v == 'if':
                self.i += 1
                self.expect('(')
                test = self.expression()
                self.expect(')')
                cons = self.statement()
                alt = self.statement() if self.eat('else') else None
                return ('if', test, cons, alt)
            if v == 'for':
                return self.for_stmt()
            if v == 'while':
                self.i += 1
                self.expect('(')
                test = self.expression()
                self.expect(')')
                return ('while', test, self.statement())
            if v == 'do':
                self.i += 1
                body = self.statement()
                self.expect('while')
                self.expect('(')
                test = self.expression()
                self.expect(')')
                self.eat(';')
                return ('dowhile', body, test)
            if v == 'break':
                self.i += 1
                self.semi()
                return ('break',)
            if v == 'continue':
                self.i += 1
                self.semi()
                return ('continue',)
            if v == 'throw':
                self.i += 1
                arg = self.expression()
                self.semi()
                return ('throw', arg)
            if v == 'try':
                self.i += 1
                blk = self.block()
                param = handler = final = None
                if self.eat('catch'):
                    if self.eat('('):
                        param = self.ident()
                        self.expect(')')
                    handler = self.block()
                if self.eat('finally'):
                    final = self.block()
                return ('try', blk, param, handler, final)
            if v == 'switch':
                self.i += 1
                self.expect('(')
                disc = self.expression()
                self.expect(')')
                self.expect('{')
                cases = []
                while not self.at('}'):
                    if self.eat('default'):
                        test = None
                    else:
                        self.expect('case')
                        test = self.expression()
                    self.expect(':')
                    body = []
                    while not (self.at('case') or self.at('default') or self.at('}')):
                        body.append(self.statement())
                    cases.append((test, body))
                self.expect('}')
                return ('switch', disc, cases)
        e = self.expression()
        self.semi()
        return ('expr', e)

    def var_decl(self):
        kind = self.ident()
        decls = []
        while True:
            name = self.ident()
            init = self.assignment() if self.eat('=') else None
            decls.append((name, init))
            if not self.eat(','):
                break
        return ('var', kind, decls)

    def for_stmt(self):
        self.expect('for')
        self.expect('(')
        init = None
        if not self.at(';'):
            if self.peek()[1] in ('var', 'let', 'const') and self.peek()[0] == 'id':
                # for (let x in obj) / for (let x of obj)
                if self.peek(2)[0] == 'id' and self.peek(2)[1] in ('in', 'of'):
                    kind = self.ident()
                    name = self.ident()
                    mode = self.ident()
                    obj = self.expression()
                    self.expect(')')
                    return ('forin' if mode == 'in' else 'forof', kind, name, obj, self.statement())
                self.no_in = True
                init = self.var_decl()
                self.no_in = False
            else:
                self.no_in = True
                e = self.expression()
                self.no_in = False
                if self.at('in') or self.at('of'):
                    mode = self.ident()
                    if e[0] != 'id':
                        raise SyntaxError("minijs: for-in target must be an identifier")
                    obj = self.expression()
                    self.expect(')')
                    return ('forin' if mode == 'in' else 'forof', None, e[1], obj, self.statement())
                init = ('expr', e)
        self.expect(';')
        test = None if self.at(';') else self.expression()
        self.expect(';')
        update = None if self.at(')') else self.expression()
        self.expect(')')
        return ('for', init, test, update, self.statement())

    def function(self, is_decl=False):
        is_async = self.eat('async')
        self.expect('function')
        if self.at('*'):
            raise SyntaxError("minijs: generators are not supported")
        name = None
        if self.peek()[0] == 'id' and not self.at('('):
            name = self.ident()
        params = self.params()
        body = self.block()
        return ('fn', name, params, body, False, is_async, False)

    def params(self):
        self.expect('(')
        out = []
        while not self.at(')'):
            rest = self.eat('...')
            if self.at('{') or self.at('['):
                raise SyntaxError("minijs: destructuring parameters are not supported")
            name = self.ident()
            default = self.assignment() if self.eat('=') else None
            out.append((name, default, rest))
            if not self.eat(','):
                break
        self.expect(')')
        return out

    # -- expressions
    def expression(self):
        e = self.assignment()
        if self.at(','):
            items = [e]
            while self.eat(','):
                items.append(self.assignment())
            return ('seq', items)
        return e

    def _arrow_ahead(self):
        """At '(': is this the parameter list of an arrow function?"""
        depth, k = 0, 0
        while True:
            tk = self.peek(k)
            if tk[0] == 'eof':
                return False
            if tk[0] == 'punc':
                if tk[1] in '([{':
                    depth += 1
                elif tk[1] in ')]}':
                    depth -= 1
                    if depth == 0:
                        nxt = self.peek(k + 1)
                        return nxt[0] == 'punc' and nxt[1] == '=>'
            k += 1

    def arrow(self, is_async=False):
        if self.at('('):
            params = self.params()
        else:
            params = [(self.ident(), None, False)]
        self.expect('=>')
        if self.at('{'):
            return ('fn', None, params, self.block(), True, is_async, False)
        saved, self.no_in = self.no_in, False
        body = self.assignment()
        self.no_in = saved
        return ('fn', None, params, body, True, is_async, True)

    def assignment(self):
        tk = self.peek()
        if tk[0] == 'id' and tk[1] not in KEYWORDS and self.peek(1)[0] == 'punc' and self.peek(1)[1] == '=>':
            return self.arrow()
        if tk[0] == 'punc' and tk[1] == '(' and self._arrow_ahead():
            return self.arrow()
        if tk[0] == 'id' and tk[1] == 'async' and not self.at('function', 1):
            nxt = self.peek(1)
            if (nxt[0] == 'punc' and nxt[1] == '(') or (nxt[0] == 'id' and self.peek(2)[1] == '=>'):
                self.i += 1
                return self.arrow(is_async=True)
        left = self.conditional()
        tk = self.peek()
        if tk[0] == 'punc' and tk[1] in _ASSIGN:
            if left[0] not in ('id', 'member', 'index'):
                raise SyntaxError(f"minijs: invalid assignment target at offset {tk[2]}")
            self.i += 1
            return ('assign', tk[1], left, self.assignment())
        return left

    def conditional(self):
        test = self.binary(0)
        if self.eat('?'):
            saved, self.no_in = self.no_in, False
            a = self.assignment()
            self.no_in = saved
            self.expect(':')
            b = self.assignment()
            return ('cond', test, a, b)
        return test

    def binary(self, minprec):
        left = self.unary()
        while True:
            tk = self.peek()
            op = tk[1]
            if tk[0] not in ('punc', 'id') or op not in _BINPREC:
                return left
            if op == 'in' and self.no_in:
                return left
            prec = _BINPREC[op]
            if prec < minprec:
                return left
            self.i += 1
            right = self.binary(prec if op == '**' else prec + 1)
            left = ('logical', op, left, right) if op in ('&&', '||') else ('bin', op, left, right)

    def unary(self):
        tk = self.peek()
        if tk[0] == 'punc':
            if tk[1] in ('!', '-', '+', '~'):
                self.i += 1
                return ('unary', tk[1], self.unary())
            if tk[1] in ('++', '--'):
                self.i += 1
                return ('update', tk[1], True, self.unary())
        elif tk[0] == 'id':
            if tk[1] in ('typeof', 'void', 'delete'):
                self.i += 1
                return ('unary', tk[1], self.unary())
            if tk[1] == 'await':
                raise SyntaxError("minijs: await is not supported (none of the evaluated modules uses it)")
        e = self.postfix()
        return e

    def postfix(self):
        e = self.call_member()
        tk = self.peek()
        if tk[0] == 'punc' and tk[1] in ('++', '--'):
            self.i += 1
            return ('update', tk[1], False, e)
        return e

    def arguments(self):
        self.expect('(')
        args = []
        while not self.at(')'):
            if self.at('...'):
                raise SyntaxError("minijs: spread arguments are not supported")
            args.append(self.assignment())
            if not self.eat(','):
                break
        self.expect(')')
        return args

    def call_member(self):
        if self.at('new'):
            self.i += 1
            callee = self.member_only()
            args = self.arguments() if self.at('(') else []
            e = ('new', callee, args)
        else:
            e = self.primary()
        while True:
            if self.eat('.'):
                tk = self.peek()
                if tk[0] != 'id':
                    raise SyntaxError(f"minijs: property name expected at offset {tk[2]}")
                self.i += 1
                e = ('member', e, tk[1])
            elif self.at('['):
                self.i += 1
                saved, self.no_in = self.no_in, False
                idx = self.expression()
                self.no_in = saved
                self.expect(']')
                e = ('index', e, idx)
            elif self.at('('):
                e = ('call', e, self.arguments())
            else:
                return e

    def member_only(self):
        if self.at('new'):
            self.i += 1
            callee = self.member_only()
            args = self.arguments() if self.at('(') else []
            e = ('new', callee, args)
        else:
            e = self.primary()
        while True:
            if self.eat('.'):
                e = ('member', e, self.ident())
            elif self.at('['):
                self.i += 1
                idx = self.expression()
                self.expect(']')
                e = ('index', e, idx)
            else:
                return e

    def primary(self):
        tk = self.peek()
        kind, val = tk[0], tk[1]
        if kind == 'num':
            self.i += 1
            return ('num', val)
        if kind == 'str':
            self.i += 1
            return ('str', val)
        if kind == 'id':
            if val == 'function' or (val == 'async' and self.at('function', 1)):
                return self.function()
            self.i += 1
            if val == 'this':
                return ('this',)
            if val == 'true':
                return ('const', True)
            if val == 'false':
                return ('const', False)
            if val == 'null':
                return ('const', None)
            if val in KEYWORDS:
                raise SyntaxError(f"minijs: unexpected keyword {val!r} at offset {tk[2]}")
            return ('id', val)
        if val == '(':
            self.i += 1
            saved, self.no_in = self.no_in, False
            e = self.expression()
            self.no_in = saved
            self.expect(')')
            return e
        if val == '[':
            self.i += 1
            items = []
            while not self.at(']'):
                if self.at(','):
                    self.i += 1
                    items.append(None)       # hole
                    continue
                if self.at('...'):
                    raise SyntaxError("minijs: spread elements are not supported")
                items.append(self.assignment())
                if not self.eat(','):
                    break
            self.expect(']')
            return ('arr', items)
        if val == '{':
            self.i += 1
            props = []
            while not self.at('}'):
                k = self.peek()
                if k[0] in ('id', 'str'):
                    key = k[1]
                elif k[0] == 'num':
                    key = k[1]
                else:
                    raise SyntaxError(f"minijs: unsupported object literal key at offset {k[2]}")
                self.i += 1
                if self.eat(':'):
                    props.append((key, self.assignment()))
                elif self.at('('):
                    params = self.params()
                    props.append((key, ('fn', key, params, self.block(), False, False, False)))
                else:
                    props.append((key, ('id', key)))
                if not self.eat(','):
                    break
            self.expect('}')
            return ('obj', props)
        raise SyntaxError(f"minijs: unexpected token {val!r} at offset {tk[2]}: {self.src[tk[2]:tk[2] + 40]!r}")
