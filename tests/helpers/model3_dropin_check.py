"""Builds the reference's NaturalSpeech2 (model3.py) with the drop-in unet1d / sampler packages installed, the five
absent third-party packages stubbed. Prints the module of every UNet1DConditionModel instance."""
import sys, types, importlib, importlib.abc, importlib.machinery, warnings, json, os
warnings.filterwarnings("ignore")
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
class Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'): raise AttributeError(name)
        m = Stub(self.__name__+'.'+name); m.__path__=[]; sys.modules[m.__name__]=m; setattr(self,name,m); return m
    def __call__(self,*a,**k): return self
class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    roots=('vocos','accelerate','matplotlib','librosa','ema_pytorch')
    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in self.roots: return importlib.machinery.ModuleSpec(name, self, is_package=True)
    def create_module(self, spec):
        m=Stub(spec.name); m.__path__=[]; return m
    def exec_module(self, m): pass
sys.meta_path.insert(0, Finder())
import diffvits_b200; diffvits_b200.install_dropin()
REF=os.environ.get('DIFFVITS_REFERENCE','/root/reference')
sys.path.append(REF)
os.chdir(REF)
import model3
cfg=json.load(open('config.json'))
import torch; torch.manual_seed(0)
m=model3.NaturalSpeech2(cfg)
print(type(m.diff_model.unet).__module__)
print([type(x).__module__ for x in m.modules() if type(x).__name__=='UNet1DConditionModel'])
