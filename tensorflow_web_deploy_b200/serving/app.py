"""py3 / werkzeug port of the reference's HTTP front-end (/root/reference/PyClassification.py:23-130).

Flask and gunicorn are not in this image (and the reference module is Python 2), so the three routes are restated on
werkzeug + Jinja2 — which is what Flask is made of — with the same URLs, methods, status codes, response bodies and
the `ClassificationAccuracy` header:

  GET|POST /            :29-55    POST: raw JPEG body -> "name1,name2,name3", header ClassificationAccuracy = sum(top-5)
                                  GET ?url= / ?file=: same from a URL / a local file; bare GET renders the page
  GET  /classify_url    :58-82    ?imageurl= -> page with the top-3 (name, score), the time and a thumbnail
  POST /classify_upload :85-113   multipart `imagefile` (PNG is re-encoded to JPEG first, :93-96) -> same page

`app.clf` is anything with the reference's classify_image contract: the in-process ImagenetClassifier or the
micro-batcher's BatcherClient.  As in the reference, a failed classification (None) surfaces as HTTP 500.
"""
from __future__ import annotations

import base64
import datetime
import io
import logging
import os
import urllib.request

import numpy as np
from jinja2 import Environment, FileSystemLoader, select_autoescape
from werkzeug.exceptions import HTTPException, InternalServerError, NotFound
from werkzeug.utils import secure_filename
from werkzeug.wrappers import Request, Response

UPLOAD_FOLDER = "/tmp/demos_uploads"                                  # PyClassification.py:24
ALLOWED_IMAGE_EXTENSIONS = {"png", "jpg", "jpe", "jpeg"}              # :25
_TEMPLATES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "templates")

log = logging.getLogger("icp.app")


def open_oriented_im(src) -> np.ndarray:
    """The display-only image of the result page (reference exifutil.py): PIL open, EXIF orientation applied,
    float32 HWC in [0,1].  Never fed to the model."""
    from PIL import Image, ImageOps
    im = Image.open(src)
    im = ImageOps.exif_transpose(im).convert("RGB")
    return np.asarray(im, np.float32) / 255.0


def embed_image_html(image: np.ndarray) -> str:
    """:116-123: 256x256 PNG thumbnail as a data URI."""
    from PIL import Image
    im = Image.fromarray((255 * image).astype("uint8")).resize((256, 256))
    buf = io.BytesIO()
    im.save(buf, format="png")
    return "data:image/png;base64," + base64.b64encode(buf.getvalue()).decode()


def allowed_file(filename: str) -> bool:
    return "." in filename and filename.rsplit(".", 1)[1] in ALLOWED_IMAGE_EXTENSIONS


class ClassifierApp:
    """WSGI application; `clf` may be set later (the reference attaches it in setup_app, :221-226)."""

    def __init__(self, clf=None):
        self.clf = clf
        self.jinja = Environment(loader=FileSystemLoader(_TEMPLATES), autoescape=select_autoescape(["html"]))
        os.makedirs(UPLOAD_FOLDER, exist_ok=True)

    # ------------------------------------------------------------------ helpers
    def render(self, **ctx) -> Response:
        return Response(self.jinja.get_template("index.html").render(**ctx), mimetype="text/html")

    def _classify(self, data: bytes):
        out = self.clf.classify_image(data)
        if out is None:                       # the reference's tuple-unpack of None raises -> HTTP 500 (:54, :216-218)
            raise InternalServerError("classification failed")
        return out

    # ------------------------------------------------------------------ routes
    def classify_index(self, request: Request) -> Response:
        string_buffer = None
        if request.method == "GET":
            url = request.args.get("url")
            if url:
                log.info("Image: %s", url)
                string_buffer = urllib.request.urlopen(url).read()
            file = request.args.get("file")
            if file:
                log.info("Image: %s", file)
                with open(file, "rb") as f:
                    string_buffer = f.read()
            if not string_buffer:
                return self.render(has_result=False)
        elif request.method == "POST":
            string_buffer = request.get_data(cache=False)
        if not string_buffer:
            return Response(status=400)
        names, probs, time_cost, accuracy = self._classify(string_buffer)
        return Response(",".join(names), 200, {"ClassificationAccuracy": str(accuracy)}, mimetype="text/html")

    def classify_url(self, request: Request) -> Response:
        imageurl = request.args.get("imageurl", "")
        try:
            data = urllib.request.urlopen(imageurl).read()
            image = open_oriented_im(io.BytesIO(data))
        except Exception as err:
            log.info("URL Image open error: %s", err)
            return self.render(has_result=True, result=(False, "Cannot open image from URL."))
        log.info("Image: %s", imageurl)
        names, probs, time_cost, accuracy = self._classify(data)
        return self.render(has_result=True, result=[True, list(zip(names, probs)), "%.3f" % time_cost],
                           imagesrc=embed_image_html(image))

    def classify_upload(self, request: Request) -> Response:
        try:
            imagefile = request.files["imagefile"]
            filename_ = str(datetime.datetime.now()).replace(" ", "_") + secure_filename(imagefile.filename)
            filename = os.path.join(UPLOAD_FOLDER, filename_)
            imagefile.save(filename)
            path, extension = os.path.splitext(filename)
            if extension == ".png":                        # the graph only has DecodeJpeg (:93-96)
                from PIL import Image
                im = Image.open(filename).convert("RGB")
                filename = "%s.jpg" % path
                im.save(filename)
            log.info("Saving to %s.", filename)
            image = open_oriented_im(filename)
        except Exception as err:
            log.info("Uploaded image open error: %s", err)
            return self.render(has_result=True, result=(False, "Cannot open uploaded image."))
        with open(filename, "rb") as f:
            names, probs, time_cost, accuracy = self._classify(f.read())
        return self.render(has_result=True, result=[True, list(zip(names, probs)), "%.3f" % time_cost],
                           imagesrc=embed_image_html(image))

    # ------------------------------------------------------------------ WSGI
    def dispatch(self, request: Request) -> Response:
        path = request.path.rstrip("/") or "/"
        if path == "/" and request.method in ("GET", "POST"):
            return self.classify_index(request)
        if path == "/classify_url" and request.method == "GET":
            return self.classify_url(request)
        if path == "/classify_upload" and request.method == "POST":
            return self.classify_upload(request)
        if path in ("/", "/classify_url", "/classify_upload"):
            return Response(status=405)
        raise NotFound()

    def __call__(self, environ, start_response):
        request = Request(environ)
        try:
            response = self.dispatch(request)
        except HTTPException as e:
            response = e
        except Exception as err:                  # anything else is a 500, as under Flask
            log.exception("unhandled error: %s", err)
            response = InternalServerError()
        return response(environ, start_response)


def setup_app(app: ClassifierApp, clf, sample_path=None):
    """:221-226: attach the classifier and warm it up on the sample picture, logging its top-3."""
    app.clf = clf
    if sample_path and os.path.exists(sample_path):
        log.info("testing sample picture...")
        with open(sample_path, "rb") as f:
            ret = clf.classify_image(f.read())
        if ret is not None:
            log.info("sample testing complete %s %s %s", ret[0][0], ret[0][1], ret[0][2])
    return app
