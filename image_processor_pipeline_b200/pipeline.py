"""ProcessingStep / ProcessingPipeline — the reference's orchestrator contract
(reference pipeline.py:15-583) around GPU-backed process functions.

What is kept verbatim (it is the plugin API): constructor arguments and
attributes, the five pairing modes and their RNG draws, the call
``process_function(*inputs, output_dirs=..., **options)``, the per-item log entries and
statuses, JSON log location, ``add_step`` chaining rules, ``run(from_step_index, only_one)``.

What differs, deliberately:
* process functions flagged ``uses_gpu`` run on a thread pool inside this process when
  ``workers > 1`` (one CUDA context; decode/encode release the GIL) instead of a
  ``ProcessPoolExecutor`` whose forked workers could not share the context;
* defects listed in SURVEY.md Appendix B are repaired (``errors_count`` typo, ``min()``
  without arguments, ``workers=None``, ``tqdm.notebook`` hard dependency).
"""
from __future__ import annotations

import concurrent.futures
import json
import random
from collections import Counter
from os import cpu_count
from pathlib import Path
from typing import Any, Callable, Dict, Iterator, List, Optional, Tuple
from warnings import warn

try:  # the reference imports tqdm.notebook, which needs ipywidgets outside Jupyter
    from tqdm.auto import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **kw):
        return it

MODES = ("one_input", "zip", "modulo", "sample", "custom")


class ProcessingStep:
    def __init__(self, name: str, process_function: Callable,
                 input_dirs=None, output_dirs=None, pairing_method: str = "one_input",
                 pairing_function: Optional[Callable[[List[List[Path]]], Iterator[Tuple]]] = None,
                 fixed_input: bool = False, root_dir=None, sample_k: Optional[int] = None,
                 save_log: bool = False, workers: Optional[int] = 1, options: Optional[Dict] = None):
        self.name = name
        self.process_function = process_function
        self.root_dir = Path(root_dir) if root_dir else None
        self.process_kwargs = options or {}
        self.sample_k = sample_k
        self.save_log = save_log
        self.input_paths: List[Path] = self._resolve_paths(input_dirs or [])
        self.output_paths: List[Path] = self._resolve_paths(output_dirs or [])
        self.fixed_input = fixed_input
        if not self.output_paths:
            raise ValueError(f"L'étape '{self.name}' doit avoir au moins un 'output_dirs' défini.")
        if pairing_method not in MODES:
            raise ValueError(f"Mode d'appariement' '{pairing_method}' invalide. Choisir parmi: {MODES}")
        if pairing_method == "custom" and not callable(pairing_function):
            raise ValueError("Une `pairing_function` valide est requise pour le mode 'custom'.")
        self.pairing_method = pairing_method
        self.pairing_function = pairing_function
        self.process_logs: List[Dict[str, Any]] = []
        max_cpus = cpu_count() or 1
        workers = 1 if workers is None else workers
        if workers > max_cpus:
            warn(f"Nombre de workers parallèles ajusté à {max_cpus} (maximum système).")
        if workers == -1:
            workers = max_cpus
        self.parallels_workers = min(workers, max_cpus)

    # ------------------------------------------------------------------ paths
    def _resolve_paths(self, dir_list) -> List[Path]:
        dir_list = dir_list if isinstance(dir_list, list) else [dir_list]
        resolved = []
        for folder in dir_list:
            if not isinstance(folder, (str, Path)):
                raise ValueError(f"un élément ne représente pas un dossier ou un chemin : {folder}")
            p = Path(folder)
            resolved.append(self.root_dir / p if (not p.is_absolute() and self.root_dir) else p)
        return resolved

    def __str__(self) -> str:
        ins = ", ".join(p.name for p in self.input_paths)
        outs = ", ".join(p.name for p in self.output_paths)
        return (f"Étape '{self.name}':\n  Entrée(s) : [{ins}] (Mode: {self.pairing_method})\n"
                f"  Sortie(s) : [{outs}]\n  Options   : {self.process_kwargs}")

    def _get_files_from_inputs(self) -> List[List[Path]]:
        if not self.input_paths:
            raise ValueError(f"{self.name} : Aucun dossier d'entrée défini.")
        print(f"Info [{self.name}]: Récupération des chemins de fichiers d'entrée...")
        lists = []
        for input_dir in self.input_paths:
            if not input_dir.is_dir():
                raise FileNotFoundError(f"Le dossier d'entrée spécifié n'existe pas: '{input_dir}' pour l'étape '{self.name}'")
            try:
                files = sorted(f for f in input_dir.iterdir() if f.is_file())
            except Exception as e:
                raise IOError(f"Échec de l'inventaire du dossier {input_dir}") from e
            print(f"  '{input_dir.name}' : {len(files)} fichiers trouvés.")
            lists.append(files)
        return lists

    # ------------------------------------------------------------------ pairing (reference pipeline.py:148-235)
    def _generate_processing_inputs(self, input_file_lists: List[List[Path]]) -> Iterator[Tuple]:
        if not all(input_file_lists):
            empty = [str(self.input_paths[i]) for i, lst in enumerate(input_file_lists) if not lst]
            raise FileNotFoundError(f"Aucun fichier trouvé dans les dossiers d'entrée {empty} pour l'étape '{self.name}'.")
        if self.sample_k and isinstance(self.sample_k, int):
            ids = random.sample(range(len(input_file_lists[0])), self.sample_k)
            input_file_lists = [[lst[i] for i in ids] for lst in input_file_lists]
        mode = self.pairing_method
        if mode == "one_input":
            for f in input_file_lists[0]:
                yield (f,)
        elif mode == "zip":
            if len(input_file_lists) < 2:
                raise ValueError("Le mode 'zip' requiert au moins 2 dossiers d'entrée.")
            yield from zip(*input_file_lists)
        elif mode == "modulo":
            if len(input_file_lists) != 2:
                raise ValueError("Le mode 'modulo' requiert exactement 2 dossiers d'entrée.")
            first, second = input_file_lists
            random.shuffle(second)
            for i, p in enumerate(first):
                yield (p, second[i % len(second)])
        elif mode == "sample":
            files = input_file_lists[0]
            blur = set(random.sample(files, int(len(files) * 0.3)))
            do_blur = [f in blur for f in files]
            rgb = set(random.sample(files, int(len(files) * 0.3)))
            do_rgb = [f in rgb for f in files]
            yield from zip(files, do_blur, do_rgb)
        elif mode == "custom":
            if not self.pairing_function:
                raise ValueError("Fonction `pairing_function` manquante pour le mode 'custom'.")
            yield from self.pairing_function(input_file_lists)
        else:  # pragma: no cover
            raise NotImplementedError(f"Mode d'appariement' '{mode}' non implémentée.")

    # ------------------------------------------------------------------ run
    def run(self):
        self.process_logs = []
        print(f"--- Exécution Étape : {self.name} ---")
        for out in self.output_paths:
            try:
                out.mkdir(parents=True, exist_ok=True)
            except IOError as ioe:
                raise IOError(f"Impossible de créer le dossier de sortie '{out}': {ioe}") from ioe
        try:
            lists = self._get_files_from_inputs()
        except (FileNotFoundError, ValueError, IOError) as e:
            print(f"Erreur [{self.name}]: Condition préalable non remplie pour démarrer l'étape. {e}")
            return
        iterator = self._generate_processing_inputs(lists)
        total = None
        if self.pairing_method in ("one_input", "modulo", "sample"):
            total = len(lists[0])
        elif self.pairing_method == "zip":
            total = min(len(lst) for lst in lists if lst) if any(lists) else None
        if self.sample_k and isinstance(self.sample_k, int):
            total = self.sample_k
        try:
            ok, bad = self._processing_loop(iterator, total)
        except (FileNotFoundError, ValueError, NotImplementedError) as e:
            print(f"Erreur [{self.name}]: Impossible de générer les arguments pour le mode '{self.pairing_method}'. {e}")
            return
        counts = Counter(log.get("status") for log in self.process_logs)
        if self.process_logs and self.save_log:
            self._save_process_logs_to_json()
        print(f"--- Étape {self.name} terminée --- {dict(counts)}")
        print(f"  {ok} éléments traités avec succès (fichiers de sortie générés).")
        if bad > 0:
            print(f"  {bad} erreur(s) ou traitement(s) sans retour.")

    def _call(self, args: Tuple):
        return self.process_function(*args, output_dirs=self.output_paths, **self.process_kwargs)

    def _processing_loop(self, argument_iterator, total_items) -> Tuple[int, int]:
        ok = bad = 0
        if not self.parallels_workers or 0 <= self.parallels_workers <= 1:
            for args in tqdm(argument_iterator, desc=self.name, total=total_items, unit="item", leave=True, smoothing=0):
                entry = {"inputs": list(args), "outputs": None, "status": "Pending", "error_message": None}
                try:
                    if self._build_log(entry, self._call(args)):
                        ok += 1
                    else:
                        bad += 1
                except Exception as e_proc:
                    print(f"\nErreur [{self.name}]: Échec traitement de {args}: {e_proc}")
                    entry.update({"status": "Error", "error_message": str(e_proc)})
                    bad += 1
                self.process_logs.append(entry)
            return ok, bad

        all_args = list(argument_iterator)
        if not all_args:
            raise RuntimeError("Aucun argument à traiter après génération. Fin.")
        gpu = getattr(self.process_function, "uses_gpu", False)
        Pool = concurrent.futures.ThreadPoolExecutor if gpu else concurrent.futures.ProcessPoolExecutor
        with Pool(max_workers=self.parallels_workers) as executor:
            pending: Dict[concurrent.futures.Future, Dict[str, Any]] = {}
            for args in all_args:
                entry = {"inputs": list(args), "outputs": None, "status": "Pending Execution", "error_message": None}
                try:
                    if gpu:
                        fut = executor.submit(self._call, args)
                    else:
                        fut = executor.submit(self.process_function, *args, output_dirs=self.output_paths, **self.process_kwargs)
                    pending[fut] = entry
                except Exception as e_submit:
                    entry.update({"status": "Submission Error", "error_message": str(e_submit)})
                    self.process_logs.append(entry)
                    bad += 1
            for fut in tqdm(concurrent.futures.as_completed(pending), total=len(pending), desc=self.name, unit="item",
                            leave=True, smoothing=0):
                entry = pending[fut]
                try:
                    if self._build_log(entry, fut.result()):
                        ok += 1
                    else:
                        bad += 1
                except Exception as e_exec:
                    msg = f"Échec tâche parallèle pour {[str(p) for p in entry['inputs']]} : {e_exec}"
                    print(f"\nErreur [{self.name}]: {msg}")
                    entry.update({"status": "Error", "error_message": msg})
                    bad += 1
                self.process_logs.append(entry)
        return ok, bad

    def _build_log(self, log_entry: Dict[str, Any], saved) -> bool:
        """Path -> Success, list of Path -> Success, other truthy -> 'Type Error',
        falsy -> 'no_output' (reference pipeline.py:439-470)."""
        if saved:
            if isinstance(saved, Path):
                log_entry.update({"outputs": [saved], "status": "Success"})
                return True
            if isinstance(saved, list) and all(isinstance(p, Path) for p in saved):
                log_entry.update({"outputs": saved, "status": "Success"})
                return True
            msg = (f"Retour invalide de {getattr(self.process_function, '__name__', '?')} pour "
                   f"{[str(p) for p in log_entry['inputs']]} (type : {type(saved)}). Attendu Path, List[Path] ou None.")
            warn(msg)
            log_entry.update({"status": "Type Error", "error_message": msg})
            return False
        log_entry["status"] = "no_output"
        return False

    def _save_process_logs_to_json(self) -> None:
        if not self.output_paths or not self.process_logs:
            return
        json_path = self.output_paths[0].parent / Path(self.name).with_suffix(".json")
        try:
            with json_path.open("w", encoding="utf-8") as j:
                json.dump(self.process_logs, j, indent=4, ensure_ascii=False, cls=PathJSONEncoder)
            print(f"Info [{self.name}]: Logs sauvegardé avec succès dans {json_path}.")
        except (IOError, TypeError) as e:
            print(f"Erreur critique [{self.name}]: Impossible d'enregistrer le fichier JSON des résultats: {e}")


class ProcessingPipeline:
    def __init__(self, root_dir=None):
        self.steps: List[ProcessingStep] = []
        self.root_dir = Path(root_dir) if root_dir else None

    def add_step(self, step: ProcessingStep, position=None):
        """Append / insert a step; a step without input_dirs is chained to the previous
        step's outputs, and a following non-fixed step is re-pointed to this step's outputs
        (reference pipeline.py:508-553)."""
        if not self.steps and not step.input_paths:
            raise ValueError(f"La première étape ('{step.name}') doit avoir `input_dirs` définie.")
        if self.root_dir and not step.root_dir:
            step.root_dir = self.root_dir
            step.input_paths = step._resolve_paths(step.input_paths)
            step.output_paths = step._resolve_paths(step.output_paths)
        position = len(self.steps) if position is None else position
        if not step.input_paths:
            if position == 0:
                raise IndexError(f"Insertion en première position, impossible de déduire les dossiers d'input pour {step.name}")
            try:
                previous = self.steps[position - 1]
                nxt = self.steps[position] if position < len(self.steps) else None
                step.input_paths = previous.output_paths
                if nxt and not nxt.fixed_input:
                    nxt.input_paths = step.output_paths
            except IndexError as idx:
                raise ValueError(f"Position d'insertion invalide pour déduire les dossiers d'input de {step.name}.") from idx
        self.steps.insert(position, step)

    def run(self, from_step_index: int = 0, only_one: bool = False):
        if from_step_index < 0 or from_step_index >= len(self.steps):
            raise IndexError(f"Invalid start index {from_step_index}. Pipeline has {len(self.steps)} steps.")
        todo = [self.steps[from_step_index]] if only_one else self.steps[from_step_index:]
        for i, step in enumerate(todo, start=from_step_index):
            print(f"Running étape {i}: {step.name}")
            step.run()


class PathJSONEncoder(json.JSONEncoder):
    """Path -> str, tuple -> list (reference pipeline.py:569-583)."""

    def default(self, o: Any) -> Any:
        if isinstance(o, Path):
            return str(o)
        if isinstance(o, tuple):
            return list(o)
        return super().default(o)
