import sys; sys.path.insert(0,'.')
from ptina_b200 import scenes, worker, _native
worker.init()
ctx=_native.context()
for name in ('cornell_boxes','cornell_monkey','matball','mini_matball','mega_small','metropolis'):
    sc=getattr(scenes,name)()
    scenes.apply(worker, sc)
    t=ctx.tree
    print(name, 'n',t.n,'depth',t.depth,'trav_depth',t.trav_depth,'ploc',t.trav_ploc,'list',t.list_n)
