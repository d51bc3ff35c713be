// ICPGpu.cs -- drop-in for vtkPointCloud.ICP (BaseClass/ICP.cs:8-308): go_hell_ICP keeps its signature and
// mutates R (Matrix 3x3) and T (Matrix 3x1) in place (Matrix.mat is row-major, Matrix.cs:30-34).  Swap
// `new ICP()` for `new ICPGpu()` at FrmMain.cs:2685.  Source only (no .NET toolchain in the build image).
using System;
using System.Collections.Generic;

namespace vtkPointCloud
{
    class ICPGpu : IDisposable
    {
        private IntPtr ctx;
        public int itersDone; public double sseLast; public int[] orderLast;   // extras the C# discards
        public int maxIters = 0;   // 0 = unbounded like the reference (ICP.cs:180 has no cap); the VTK path uses 100 (FrmMain.cs:855)

        public ICPGpu() { ctx = DBImprovedGpu.SharedContext(); }   // one native context per application (DBImprovedGpu.Devices selects the GPUs)

        private static double[] Planar(List<Point3D> pts)
        {
            int k = pts.Count; double[] a = new double[3 * k];
            for (int i = 0; i < k; i++) { a[i] = pts[i].X; a[k + i] = pts[i].Y; a[2 * k + i] = pts[i].Z; }
            return a;
        }

        public void go_hell_ICP(List<Point3D> model, List<Point3D> data, Matrix R, Matrix T, double e)
        {
            orderLast = new int[data.Count];
            NativeMethods.Check(ctx, NativeMethods.vpc_icp_rigid(ctx, Planar(model), model.Count, Planar(data), data.Count, e, maxIters,
                R.mat, T.mat, out itersDone, out sseLast, orderLast));
        }

        // FindClosestPointSet (ICP.cs:224-250) returns the matched model points
        public List<Point3D> FindClosestPointSet(List<Point3D> model, List<Point3D> data)
        {
            int[] order = new int[data.Count];
            NativeMethods.Check(ctx, NativeMethods.vpc_closest_point_set(ctx, Planar(model), model.Count, Planar(data), data.Count, order, null));
            List<Point3D> Y = new List<Point3D>(data.Count);
            for (int i = 0; i < order.Length; i++) Y.Add(model[order[i]]);
            return Y;
        }

        public void Dispose() { ctx = IntPtr.Zero; }   // shared context: DBImprovedGpu.Shutdown() releases it
    }
}
