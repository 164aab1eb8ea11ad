"""Command-line flags with the reference's TF-style surface (core/flags.py of the reference):
DEFINE_string / DEFINE_integer / DEFINE_float / DEFINE_boolean / DEFINE_version register options on one
module-level argparse parser; the module-level FLAGS object parses lazily on first attribute access,
ignores unknown arguments (parse_known_args, core/flags.py:22), accepts `--flag`, `--flag=true`,
`--flag False` and `--noflag` for booleans (core/flags.py:81-104) and lets values be overridden by
assignment."""
import argparse

_parser = argparse.ArgumentParser(description='B200-native GPU-cluster scheduling simulator')
_TRUE = ('true', 't', '1')


class _Flags(object):
    def __init__(self):
        object.__setattr__(self, '_values', {})
        object.__setattr__(self, '_parsed', False)

    def _ensure(self, args=None):
        if not self._parsed or args is not None:
            ns, unknown = _parser.parse_known_args(args=args)
            self._values.update(vars(ns))
            object.__setattr__(self, '_parsed', True)
            return unknown
        return []

    def parse(self, args=None):
        """Explicit (re)parse, e.g. with a custom argv; returns the unknown arguments."""
        return self._ensure(args if args is not None else None) if not self._parsed or args is None else self._ensure(args)

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        self._ensure()
        try:
            return self._values[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self._ensure()
        self._values[name] = value

    def get(self, name, default=None):
        self._ensure()
        return self._values.get(name, default)


FLAGS = _Flags()


def _define(name, default, doc, kind):
    _parser.add_argument('--' + name, default=default, help=doc, type=kind)


def DEFINE_string(name, default, doc):
    _define(name, default, doc, str)


def DEFINE_integer(name, default, doc):
    _define(name, default, doc, int)


def DEFINE_float(name, default, doc):
    _define(name, default, doc, float)


def DEFINE_boolean(name, default, doc):
    _parser.add_argument('--' + name, nargs='?', const=True, default=default, help=doc,
                         type=lambda v: v.lower() in _TRUE)
    _parser.add_argument('--no' + name, action='store_false', dest=name.replace('-', '_'))


DEFINE_bool = DEFINE_boolean


def DEFINE_version(version):
    _parser.add_argument('-v', '--version', action='version', version='%(prog)s ' + version, dest='version',
                         help='display version information')


def reset_for_tests():
    """Drops parsed values (definitions stay): lets tests parse several argv in one process."""
    object.__setattr__(FLAGS, '_values', {})
    object.__setattr__(FLAGS, '_parsed', False)
