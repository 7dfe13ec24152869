"""API-lifecycle helpers with the contract of the reference's `pylibcugraph.utilities.api_tools`
(python/pylibcugraph/pylibcugraph/utilities/api_tools.py): wrap a function or class whose name starts with `EXPERIMENTAL__` so
that calling it warns (PendingDeprecationWarning) and the wrapper carries the name without the prefix.  Written from the
behaviour the reference's tests/test_utils.py checks."""
import functools
import inspect
import types
import warnings

_PREFIX = "EXPERIMENTAL__"


def _public_name(obj):
    name = obj.__name__
    return name[len(_PREFIX):] if name.startswith(_PREFIX) else name


def experimental_warning_wrapper(obj):
    """Returns a stand-in for `obj` (a function or a class) that emits a PendingDeprecationWarning when it is called."""
    if not isinstance(obj, (types.FunctionType, types.BuiltinFunctionType)) and not inspect.isclass(obj):
        raise TypeError(f"obj must be a class or a function type, got {type(obj)}")
    name = _public_name(obj)
    caller = inspect.stack()[1]
    module = inspect.getmodule(caller[0])
    namespace = module.__name__ if module is not None else "__main__"
    msg = (f"{namespace}.{name} is experimental and will change or be removed in a future release.")

    if inspect.isclass(obj):
        class _Experimental(obj):
            def __init__(self, *args, **kwargs):
                warnings.warn(msg, PendingDeprecationWarning)
                super().__init__(*args, **kwargs)

        _Experimental.__module__ = namespace
        _Experimental.__qualname__ = name
        _Experimental.__name__ = name
        return _Experimental

    @functools.wraps(obj)
    def _experimental_call(*args, **kwargs):
        warnings.warn(msg, PendingDeprecationWarning)
        return obj(*args, **kwargs)

    _experimental_call.__module__ = namespace
    _experimental_call.__qualname__ = name
    _experimental_call.__name__ = name
    return _experimental_call
