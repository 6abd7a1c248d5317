"""Poor man's pyflakes (no linters in the image): compile every module and report names that are loaded in a function
but never bound anywhere in the module, builtins or the function — catches the NameError class of slips on the CPU."""
import ast
import builtins
import sys


def check(path):
    tree = ast.parse(open(path).read(), path)
    module_names = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                module_names.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            module_names.add(node.name)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
            module_names.add(node.id)
        elif isinstance(node, ast.arg):
            module_names.add(node.arg)
        elif isinstance(node, ast.ExceptHandler) and node.name:
            module_names.add(node.name)
        elif isinstance(node, (ast.Global, ast.Nonlocal)):
            module_names.update(node.names)
    bad = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load) and node.id not in module_names:
            bad.append((node.lineno, node.id))
    return bad


if __name__ == "__main__":
    rc = 0
    for p in sys.argv[1:]:
        for line, name in check(p):
            print(f"{p}:{line}: undefined name {name!r}")
            rc = 1
    sys.exit(rc)
