"""Extract every name the reference imports from a module the overlay shadows.

    python tests/golden/make_overlay_imports.py          # needs /root/reference

AST-walks /root/reference/lib and /root/reference/run and writes
tests/golden/overlay_imports.json: [{"module", "name", "importers": [file:line, ...]}].
tests/test_dropin_overlay.py re-derives the list when the reference is present (and compares),
and checks on every machine that each name resolves with the overlay in front.
"""
import ast
import json
import os
import warnings

REF = '/root/reference'
SHADOWED = ('core.inference', 'multiviews.cameras', 'multiviews.triangulate', 'multiviews.pictorial',
            'multiviews.body', 'utils.transforms')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'overlay_imports.json')


def extract(ref=REF):
    found = {}
    for top in ('lib', 'run'):
        for dirpath, _, files in os.walk(os.path.join(ref, top)):
            for fn in sorted(files):
                if not fn.endswith('.py'):
                    continue
                path = os.path.join(dirpath, fn)
                rel = os.path.relpath(path, ref)
                try:
                    with warnings.catch_warnings():
                        warnings.simplefilter('ignore')          # the reference has '\\m' style escapes
                        tree = ast.parse(open(path, encoding='utf-8', errors='replace').read())
                except SyntaxError:
                    continue
                aliases = {}
                for node in ast.walk(tree):
                    if isinstance(node, ast.ImportFrom) and node.module in SHADOWED and node.level == 0:
                        for a in node.names:
                            found.setdefault((node.module, a.name), []).append('%s:%d' % (rel, node.lineno))
                    elif isinstance(node, ast.Import):
                        for a in node.names:
                            if a.name in SHADOWED and a.asname:
                                aliases[a.asname] = a.name
                for node in ast.walk(tree):      # `import multiviews.cameras as cameras; cameras.project_pose(...)`
                    if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) \
                            and node.value.id in aliases:
                        found.setdefault((aliases[node.value.id], node.attr), []).append(
                            '%s:%d' % (rel, node.lineno))
    return [{'module': m, 'name': n, 'importers': sorted(set(v))} for (m, n), v in sorted(found.items())]


if __name__ == '__main__':
    rows = extract()
    with open(OUT, 'w') as f:
        json.dump(rows, f, indent=1)
    print('%d imported names -> %s' % (len(rows), OUT))
