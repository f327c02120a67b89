#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <omp.h>
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}

/* ---- sinf restatement (double evaluation), FMA contraction selectable ---- */
#ifndef USE_FMA
#define USE_FMA 1
#endif
#if USE_FMA
#define MADD(a,b,c) fma((a),(b),(c))
#else
static inline double MADD(double a,double b,double c){volatile double t=a*b;return t+c;}
#endif
static const double HPI_INV=0x1.45F306DC9C883p+23, HPI=0x1.921FB54442D18p0;
static const double C0=0x1p0,C1=-0x1.ffffffd0c621cp-2,C2=0x1.55553e1068f19p-5,C3=-0x1.6c087e89a359dp-10,C4=0x1.99343027bf8c3p-16;
static const double S1=-0x1.555545995a603p-3,S2=0x1.1107605230bc4p-7,S3=-0x1.994eb3774cf24p-13;
static inline float sinf_poly(double x,double x2,int neg,int n){
  if((n&1)==0){
    double x3=x*x2; double s1=MADD(x2,S3,S2); double x7=x3*x2; double s=MADD(x3,S1,x); return (float)MADD(x7,s1,s);
  } else {
    double sg = neg?-1.0:1.0;
    double c0=sg*C0,c1=sg*C1,c2=sg*C2,c3=sg*C3,c4=sg*C4;
    double x4=x2*x2; double cc2=MADD(x2,c4,c3); double cc1=MADD(x2,c1,c0); double x6=x4*x2; double c=MADD(x4,c2,cc1); return (float)MADD(x6,cc2,c);
  }
}
static float my_sinf(float y){
  double x=y; uint32_t top=(f2u(y)>>20)&0x7ff;
  if(top < ((f2u(0x1.921FB6p-1f)>>20)&0x7ff)){
    double s=x*x;
    if(top < ((f2u(0x1p-12f)>>20)&0x7ff)) return y;
    return sinf_poly(x,s,0,0);
  } else if (top < ((f2u(120.0f)>>20)&0x7ff)){
    double r=x*HPI_INV; int n=((int32_t)r+0x800000)>>24;
    double xr=MADD(-(double)n,HPI,x);
    static const double sign[4]={1.0,-1.0,-1.0,1.0};
    double s=sign[n&3];
    return sinf_poly(xr*s, xr*xr, (n&2)!=0, n);
  }
  return sinf(y);
}

/* ---- atanf / atan2f restatement (float evaluation, no contraction) ---- */
static const float atanhi[]={4.6364760399e-01f,7.8539812565e-01f,9.8279368877e-01f,1.5707962513e+00f};
static const float atanlo[]={5.0121582440e-09f,3.7748947079e-08f,3.4473217170e-08f,7.5497894159e-08f};
static const float aT[]={3.3333334327e-01f,-2.0000000298e-01f,1.4285714924e-01f,-1.1111110449e-01f,9.0908870101e-02f,-7.6918758452e-02f,6.6610731184e-02f,-5.8335702866e-02f,4.9768779427e-02f,-3.6531571299e-02f,1.6285819933e-02f};
#pragma STDC FP_CONTRACT OFF
static float my_atanf(float x){
  float w,s1,s2,z; int32_t ix,hx,id; hx=(int32_t)f2u(x); ix=hx&0x7fffffff;
  if(ix>=0x4c000000){ if(ix>0x7f800000) return x+x; if(hx>0) return atanhi[3]+atanlo[3]; else return -atanhi[3]-atanlo[3]; }
  if(ix<0x3ee00000){ if(ix<0x31000000){ return x; } id=-1; }
  else { x=fabsf(x);
    if(ix<0x3f980000){ if(ix<0x3f300000){id=0;x=(2.0f*x-1.0f)/(2.0f+x);} else {id=1;x=(x-1.0f)/(x+1.0f);} }
    else { if(ix<0x401c0000){id=2;x=(x-1.5f)/(1.0f+1.5f*x);} else {id=3;x=-1.0f/x;} } }
  z=x*x; w=z*z;
  s1=z*(aT[0]+w*(aT[2]+w*(aT[4]+w*(aT[6]+w*(aT[8]+w*aT[10])))));
  s2=w*(aT[1]+w*(aT[3]+w*(aT[5]+w*(aT[7]+w*aT[9]))));
  if(id<0) return x-x*(s1+s2);
  z=atanhi[id]-((x*(s1+s2)-atanlo[id])-x);
  return (hx<0)?-z:z;
}
static const float pi_o_2=1.5707963705e+00f, pi=3.1415927410e+00f, pi_lo=-8.7422776573e-08f, tiny=1.0e-30f;
static float my_atan2f(float y,float x){
  float z; int32_t k,m,hx,hy,ix,iy; hx=(int32_t)f2u(x); ix=hx&0x7fffffff; hy=(int32_t)f2u(y); iy=hy&0x7fffffff;
  if(ix>0x7f800000||iy>0x7f800000) return x+y;
  if(hx==0x3f800000) return my_atanf(y);
  m=((hy>>31)&1)|((hx>>30)&2);
  if(iy==0){ switch(m){case 0:case 1:return y;case 2:return pi+tiny;default:return -pi-tiny;} }
  if(ix==0) return (hy<0)?-pi_o_2-tiny:pi_o_2+tiny;
  k=(iy-ix)>>23;
  if(k>60) z=pi_o_2+0.5f*pi_lo; else if(hx<0&&k<-60) z=0.0f; else z=my_atanf(fabsf(y/x));
  switch(m){case 0:return z; case 1:return u2f(f2u(z)^0x80000000u); case 2:return pi-(z-pi_lo); default:return (z-pi_lo)-pi;}
}
int main(int argc,char**argv){
  /* sinf: all floats with |x| <= 8 */
  long bad=0; uint32_t hi=f2u(8.0f);
  #pragma omp parallel for reduction(+:bad) schedule(static)
  for(uint32_t u=0;u<=hi;u++){ for(int sgn=0;sgn<2;sgn++){ float x=u2f(u|((uint32_t)sgn<<31)); if(f2u(my_sinf(x))!=f2u(sinf(x))){ if(bad<5) printf("sinf mismatch x=%a mine=%a libm=%a\n",x,my_sinf(x),sinf(x)); bad++; } } }
  printf("sinf mismatches: %ld of %u\n",bad,2*(hi+1));
  /* atanf: all non-negative floats up to inf + negatives by symmetry sample */
  long bad2=0;
  #pragma omp parallel for reduction(+:bad2) schedule(static)
  for(uint32_t u=0;u<0x7f800000u;u++){ float x=u2f(u); if(f2u(my_atanf(x))!=f2u(atanf(x))){ if(bad2<5) printf("atanf mismatch x=%a mine=%a libm=%a\n",x,my_atanf(x),atanf(x)); bad2++; } float nx=-x; if(f2u(my_atanf(nx))!=f2u(atanf(nx))) bad2++; }
  printf("atanf mismatches: %ld\n",bad2);
  /* atan2f on integer-valued moments */
  long bad3=0; 
  #pragma omp parallel for reduction(+:bad3) schedule(static)
  for(int i=0;i<200000000;i++){ uint64_t s=0x9E3779B97F4A7C15ull*(uint64_t)(i+1); s^=s>>29; s*=0xBF58476D1CE4E5B9ull; s^=s>>32;
    int a=(int)(s%2400001)-1200000, b=(int)((s>>32)%2400001)-1200000; if(i%97==0)a=0; if(i%89==0)b=0; if(i%83==0){a=(int)(s%2001)-1000;b=(int)((s>>20)%2001)-1000;}
    float y=(float)a,x=(float)b; if(f2u(my_atan2f(y,x))!=f2u(atan2f(y,x))){ if(bad3<5) printf("atan2f mismatch y=%d x=%d mine=%a libm=%a\n",a,b,my_atan2f(y,x),atan2f(y,x)); bad3++; } }
  printf("atan2f mismatches: %ld\n",bad3);
  return 0;
}
